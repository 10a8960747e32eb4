#!/usr/bin/env python
"""
bench.py -- MLUPS of the fused lattice Boltzmann time step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

One JSON line on rank 0.  A "step" is one `one_time_step` of the whole lattice:
ghost update -> boundary kernels -> fused pull stream+collide -> swap.

Default workload = BASELINE config 4: D3Q19 MRT lid-driven cavity 512^3, fp64, the
configuration the north-star target is quoted on; it fits one B200 (41 GB for F+Fnew).
For N > 1 the SAME global lattice is cut into N x-slabs (strong scaling), one process
per GPU, halo planes stored by the fused kernel into the neighbours' memory over NVLink.

The simulation is built through the north-star interface -- the UNCHANGED
`pylbm.Simulation(dico)` of the reference installed in oracle/_ref (tools/make_ref.sh) with
`generator='cuda'` registered by `pylbm_b200.plugin` -- when that installation is present
(`run.api` says which); otherwise through this package's own front end `pylbm_b200.Simulation`.

  value        MLUPS = K * interior cells / device time (max over ranks), state resident in HBM,
               K steps enqueued by `sol.run(K)`; `stepwise` = the same K steps as K Python calls
               of `sol.one_time_step()` (CUDA events on the simulation's stream)
  roofline     fused kernel only: 2*Q*sizeof bytes per cell * cells / (CUDA-event time per launch)
               against MEASURED_PEAKS.json hbm_gbs
  e2e          same metric through the public API with HOST buffers: pinned host F -> device,
               K x sol.one_time_step(), conserved moments -> host (copies inside the timed region)
  cpu_baseline the REFERENCE's own Cython generator (oracle/_ref, 1 rank x 1 core: its generated
               code is single-threaded and the image has no MPI) on a bounded sample of the
               workload; `cpu_port` = the oracle's OpenMP C restatement on all host cores
  parity       N > 1 only: two small cases run on the same ranks (fused peer halo, graph pairs, an
               outside write, boundary_condition()) and compared with the oracle on rank 0
  also         the other BASELINE configs measured in the same run (short)

--impl reference times the reference's CPU implementation alone, same metric/config.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (case, kwargs, description)
    "d3q19_lid_512": ("lid_cavity_d3q19", dict(n=512), "D3Q19 MRT lid-driven cavity 512^3 fp64 (BASELINE config 4)"),
    "d3q19_lid_256": ("lid_cavity_d3q19", dict(n=256), "D3Q19 MRT lid-driven cavity 256^3 fp64"),
    "d2q9_karman_4096x1024": ("karman_d2q9", dict(nx=4096, ny=1024),
                              "D2Q9 Karman vortex street 4096x1024, circular obstacle (BASELINE config 2)"),
    "d2q4x3_shallow_water_4096": ("shallow_water_d2q4", dict(n=4096),
                                  "D2Q4^3 vectorial shallow water 4096^2 (BASELINE config 3)"),
    "d2q9_lid_256": ("lid_cavity_d2q9", dict(n=256), "D2Q9 lid-driven cavity 256^2 (BASELINE config 1)"),
    "d3q27_channel_512x256x256": ("channel_sphere_d3q27", dict(nx=512, ny=256, nz=256),
                                  "D3Q27 channel with sphere 512x256x256 (BASELINE config 5, one slab)"),
    # weak scaling: nx is multiplied by the number of GPUs (128 x 1024 x 1024 cells per GPU -> 1024^3 on 8)
    "d3q27_channel_weak": ("channel_sphere_d3q27", dict(nx=128, ny=1024, nz=1024),
                           "D3Q27 channel with sphere, weak scaling 128x1024x1024 per GPU (BASELINE config 5)"),
}
WEAK = {"d3q27_channel_weak"}
DEFAULT_WORKLOAD = "d3q19_lid_512"
# (workload, steps) measured beside the headline in the same run
ALSO_1GPU = [("d2q9_lid_256", 2000), ("d2q9_karman_4096x1024", 400), ("d2q4x3_shallow_water_4096", 100),
             ("d3q27_channel_512x256x256", 40)]
ALSO_NGPU = [("d3q27_channel_weak", 30)]
ALSO_WITH_REFERENCE = {"d2q9_lid_256"}      # BASELINE config 1 is the reference's own CPU-runnable case
# size of the reference's CPU sample per case: the reference builds dense [unvtot, nx, ny, nz] arrays
# (domain.py:285-293) and runs on one core, so the big configurations are sampled at reduced size
REFERENCE_SAMPLE = {
    "lid_cavity_d3q19": dict(n=128), "lid_cavity_d2q9": dict(n=256), "karman_d2q9": dict(nx=1024, ny=256),
    "shallow_water_d2q4": dict(n=1024), "channel_sphere_d3q27": dict(nx=128, ny=64, nz=64),
}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = max(smax, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown CPU"


# ---------------------------------------------------------------------------
# the reference (oracle/_ref)
# ---------------------------------------------------------------------------
def load_reference():
    """import the unmodified reference installed by tools/make_ref.sh, or return None."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "pylbm")):
        return None
    for p in (ref, os.path.join(ref, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging

    import pylbm

    logging.getLogger("pylbm").setLevel(logging.ERROR)
    return pylbm


def reference_cython_run(case, steps, warmup):
    """MLUPS of the reference's own `generator='cython'` path (1 rank x 1 core) on the reference-sized
    sample of the workload; the generated Cython module is built in a scratch directory and the build
    is not timed (monitoring.py:104-111 counts steps * interior cells / seconds the same way)."""
    import tempfile

    import numpy as np
    from pylbm_b200 import cases

    pylbm = load_reference()
    if pylbm is None:
        return None
    name, kw = case
    sample = dict(kw)
    limit = REFERENCE_SAMPLE.get(name, {})
    for key, cap in limit.items():
        if key in sample:
            sample[key] = min(sample[key], cap)
    dico = cases.CASES[name](mod=pylbm, generator="cython", **sample)
    scratch = tempfile.mkdtemp(prefix="pylbm_ref_")
    dico["codegen_option"] = {"directory": scratch}
    t0 = time.perf_counter()
    sol = pylbm.Simulation(dico)
    t_build = time.perf_counter() - t0
    for _ in range(warmup):
        sol.one_time_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sol.one_time_step()
    dt = time.perf_counter() - t0
    cells = float(np.prod(sol.domain.shape_in))
    return {
        "value": cells * steps / dt / 1e6, "unit": "MLUPS", "cores": 1, "kind": "reference",
        "sample": "%s %s: pylbm %s generator='cython' (unmodified reference, oracle/_ref), 1 rank x 1 core, "
                  "%d steps after %d warm-up in %.1f s (construction + Cython build %.0f s not timed), %s"
                  % (name, sample, pylbm.__version__, steps, warmup, dt, t_build, cpu_model()),
        "ms_per_step": dt / steps * 1e3, "ranks": 1,
    }


def cpu_port_run(case, steps, warmup, budget_s=60.0, threads=None):
    """MLUPS of the oracle's OpenMP C restatement of the reference path on a bounded sample."""
    import numpy as np
    from pylbm_b200 import cases
    from oracle.lbm_oracle import OracleSimulation

    threads = threads or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    name, kw = case

    def sized(n):
        k = dict(kw)
        if "n" in k:
            k["n"] = n
        else:
            scale = n / max(k.get("ny", n), 1)
            for key in ("nx", "ny", "nz"):
                if key in k:
                    k[key] = max(8, int(round(k[key] * scale / 8)) * 8)
        return k

    dim3 = name in ("lid_cavity_d3q19", "channel_sphere_d3q27")
    probe_n = 48 if dim3 else 256
    sim = OracleSimulation(cases.CASES[name](**sized(probe_n)), openmp=True, threads=threads)
    sim.one_time_step()
    t0 = time.perf_counter()
    for _ in range(3):
        sim.one_time_step()
    rate = 3 * float(np.prod(sim.domain.shape_in)) / (time.perf_counter() - t0)  # cells/s
    total_steps = steps + warmup
    cells_budget = rate * budget_s / max(total_steps, 1)
    full_n = kw.get("n", kw.get("ny", 256))
    candidates = [n for n in (32, 48, 64, 96, 128, 192, 256, 384, 512, 1024, 2048, 4096) if n <= full_n]
    chosen = candidates[0]
    for n in candidates:
        k = sized(n)
        cells = 1
        for key in ("nx", "ny", "nz"):
            if key in k:
                cells *= k[key]
        if "n" in k:
            cells = k["n"] ** (3 if dim3 else 2)
        if cells <= cells_budget:
            chosen = n
    kws = sized(chosen)
    sim = OracleSimulation(cases.CASES[name](**kws), openmp=True, threads=threads)
    for _ in range(warmup):
        sim.one_time_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.one_time_step()
    dt = time.perf_counter() - t0
    cells = float(np.prod(sim.domain.shape_in))
    return {
        "value": cells * steps / dt / 1e6, "unit": "MLUPS", "cores": threads, "kind": "port",
        "sample": "%s %s, %d steps after %d warm-up, %.1f s, OpenMP C restatement of the reference "
                  "Cython one_time_step (oracle/lbm_oracle.py), %s" % (name, kws, steps, warmup, dt, cpu_model()),
        "ms_per_step": dt / steps * 1e3,
    }


# ---------------------------------------------------------------------------
# device arm
# ---------------------------------------------------------------------------
class Context:
    """process-wide state of one bench run (ranks, library handles, command line)."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.nccl_id = None
        self.pylbm = None
        self.api = "pylbm_b200.Simulation"

    def init(self):
        from pylbm_b200 import runtime as rt

        self.rt = rt
        self.lib = rt.lib()
        if self.world > 1:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(self.local_rank)
            rt.check(self.lib.lbm_set_device(self.local_rank), "lbm_set_device")
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
            buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                raw = (ctypes.c_char * 128)()
                rt.check(self.lib.lbm_comm_unique_id(raw), "lbm_comm_unique_id")
                buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
            dist.broadcast(buf, 0)
            self.nccl_id = bytes(buf.cpu().numpy().tobytes())
        else:
            rt.check(self.lib.lbm_set_device(self.local_rank), "lbm_set_device")
        if self.args.api in ("auto", "pylbm"):
            self.pylbm = load_reference()
            if self.pylbm is None and self.args.api == "pylbm":
                raise SystemExit("--api pylbm: oracle/_ref is not installed (run tools/make_ref.sh)")
        if self.pylbm is not None:
            from pylbm_b200 import plugin

            plugin.register()
            self.api = "pylbm.Simulation (reference %s from oracle/_ref, generator='cuda' via pylbm_b200.plugin)" \
                       % self.pylbm.__version__

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def gather_fn(self):
        if self.world > 1 and self.args.halo == "peer":
            def gather(blob):
                out = [None] * self.world
                self.dist.all_gather_object(out, blob)
                return out
            return gather
        return None

    def max_over_ranks(self, *values):
        if self.dist is None:
            return list(values)
        import torch

        t = torch.tensor(list(values), device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(self, *values):
        if self.dist is None:
            return list(values)
        import torch

        t = torch.tensor(list(values), device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def simulation(self, case_name, case_kw, dtype="float64", compute=None):
        """the simulation object of one workload, through the north-star API when it is installed."""
        from pylbm_b200 import cases

        slab = (self.rank, self.world) if self.world > 1 else None
        if self.pylbm is not None:
            from pylbm_b200 import plugin

            plugin.configure(slab=slab, nccl_id=self.nccl_id, gather=self.gather_fn(), halo=self.args.halo,
                             compute_dtype=compute)
            dico = cases.CASES[case_name](mod=self.pylbm, generator="cuda", **case_kw)
            if self.args.in_place:
                dico["cuda_option"] = {"in_place": True}
            return self.pylbm.Simulation(dico, dtype=dtype)
        import pylbm_b200

        dico = cases.CASES[case_name](**case_kw)
        return pylbm_b200.Simulation(dico, dtype=dtype, slab=slab, nccl_id=self.nccl_id, gather=self.gather_fn(),
                                     compute_dtype=compute, in_place=self.args.in_place)


def timed_run(ctx, sim, steps, stepwise=False, graph=True):
    """device time (ms, max over ranks) of `steps` time steps: one `sol.run(steps)` call, or `steps`
    Python calls of `sol.one_time_step()`; CUDA events on the simulation's stream."""
    lib, rt = ctx.lib, ctx.rt
    ctx.barrier()
    sim.synchronize()
    launches0 = lib.lbm_sim_launch_count(sim._handle)
    rt.check(lib.lbm_sim_timer_start(sim._handle), "timer_start")
    if stepwise:
        for _ in range(steps):
            sim.one_time_step()
    else:
        sim.run(steps, graph=graph)
    ms = ctypes.c_float()
    rt.check(lib.lbm_sim_timer_stop(sim._handle, ctypes.byref(ms)), "timer_stop")
    sim.synchronize()
    ctx.barrier()
    launches = lib.lbm_sim_launch_count(sim._handle) - launches0
    (elapsed,) = ctx.max_over_ranks(float(ms.value))
    return elapsed, int(launches)


def fused_kernel_roofline(ctx, sim, steps, step_ms, workload, dtype, compute):
    """per-launch CUDA-event time of the fused kernel -> achieved algorithmic bandwidth."""
    import numpy as np

    lib, rt = ctx.lib, ctx.rt
    Q, itemsize = sim.container.nv, sim.container.F.itemsize
    local_cells = float(np.prod(sim.domain.shape_in))
    rt.check(lib.lbm_sim_profile(sim._handle, 1), "profile")
    sim.run(steps, graph=False)
    fused_ms, nl = ctypes.c_double(), ctypes.c_int64()
    rt.check(lib.lbm_sim_profile_read(sim._handle, ctypes.byref(fused_ms), ctypes.byref(nl)), "profile_read")
    rt.check(lib.lbm_sim_profile(sim._handle, 0), "profile")
    per_launch_ms = fused_ms.value / max(nl.value, 1)
    bytes_per_launch = 2.0 * Q * itemsize * local_cells
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic, traffic_src = None, None
    variant = "" if dtype == "float64" else ("_f32" if compute == "float32" else "_f32storage")
    tpath = os.path.join(ROOT, "profiles", "traffic_%s%s.json" % (workload, variant))
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            traffic_src = "profiles/%s (ncu --set full capture of this kernel, not measured in this run)" \
                          % os.path.basename(tpath)
        except Exception:
            traffic = None
    return {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "kernel": "lbmk_kernel_one_time_step",
        "launch_ms": per_launch_ms,
        "kernel_share_of_step": per_launch_ms / step_ms if ctx.world == 1 else None,
        "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
        "roofline_mlups_per_gpu": peak * 1e9 / (2.0 * Q * itemsize) / 1e6,
    }


def measure_also(ctx, workload, steps):
    """one of the other BASELINE configs, short: value, stepwise value, roofline of the fused kernel."""
    import numpy as np

    case_name, case_kw, description = WORKLOADS[workload]
    if workload in WEAK:
        case_kw = dict(case_kw, nx=case_kw["nx"] * ctx.world)
    t0 = time.time()
    sim = ctx.simulation(case_name, case_kw)
    setup = time.time() - t0
    cells = float(np.prod(sim.domain.global_size))
    sim.run(5)
    sim.synchronize()
    ms, launches = timed_run(ctx, sim, steps)
    ms_sw, _ = timed_run(ctx, sim, steps, stepwise=True)
    roof = fused_kernel_roofline(ctx, sim, min(steps, 50), ms / steps, workload, "float64", None)
    out = {
        "workload": description, "n_gpus": ctx.world, "steps": steps, "value": cells * steps / (ms * 1e-3) / 1e6,
        "unit": "MLUPS", "ms_per_step": ms / steps, "launches_per_step": launches / steps,
        "stepwise": {"value": cells * steps / (ms_sw * 1e-3) / 1e6, "ms_per_step": ms_sw / steps},
        "roofline_frac_kernel": roof["frac"], "kernel_launch_ms": roof["launch_ms"],
        "frac_of_roofline": cells * steps / (ms * 1e-3) / 1e6 / (ctx.world * roof["roofline_mlups_per_gpu"]),
        "scaling": "weak" if workload in WEAK else "strong", "setup_s": round(setup, 2),
    }
    del sim
    if workload in ALSO_WITH_REFERENCE and ctx.world == 1 and not ctx.args.no_cpu_baseline:
        # the configuration the reference runs as it is on a CPU: its Cython generator beside the number
        try:
            res = reference_cython_run((case_name, case_kw), 20, 3)
            if res is not None:
                out["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:
            out["cpu_baseline"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    return out


PARITY_CASES = [("lid_cavity_d3q19", dict(n=64)), ("karman_d2q9", dict(nx=128, ny=32))]


def multi_gpu_parity(ctx):
    """N > 1: the slab path (fused peer halo or NCCL) against the oracle on small cases, same ranks.
    30 steps = single steps, a `run` of graph pairs, an outside write of F and a stand-alone
    boundary_condition(); conserved moments gathered on rank 0, max|delta| / max|ref| <= 1e-12."""
    import numpy as np
    from pylbm_b200 import cases

    worst, detail = 0.0, []
    for name, kw in PARITY_CASES:
        sim = ctx.simulation(name, kw)
        for _ in range(5):
            sim.one_time_step()
        sim.run(12)
        sim.F_halo[1] = sim.F_halo[1]          # outside write: next step refreshes the ghosts itself
        sim.one_time_step()
        sim.boundary_condition()
        sim.run(11)
        sim.one_time_step()
        assert sim.nt == 30
        fields = {str(k): np.ascontiguousarray(sim.m[k]) for k in sim.scheme.consm}
        parts = [None] * ctx.world
        ctx.dist.all_gather_object(parts, fields)
        del sim
        if ctx.rank == 0:
            from oracle.lbm_oracle import OracleSimulation

            ora = OracleSimulation(cases.CASES[name](**kw))
            for _ in range(30):
                ora.one_time_step()
            fluid = ora.domain.in_or_out[tuple(slice(v, -v) for v in ora.domain.stencil.vmax)] == ora.domain.valin
            for key in ora.scheme.consm:
                got = np.concatenate([p[str(key)] for p in parts], axis=0)
                want = ora.m[key]
                err = float(np.abs(got[fluid] - want[fluid]).max() / np.abs(want[fluid]).max())
                worst = max(worst, err)
                detail.append({"case": name, "moment": str(key), "rel_err": err})
    return {"max_rel_err": worst, "ok": bool(worst <= 1e-12), "tolerance": 1e-12, "steps": 30,
            "cases": ["%s %s" % c for c in PARITY_CASES], "checker": "oracle/lbm_oracle.py on rank 0",
            "halo": ctx.args.halo, "worst": sorted(detail, key=lambda d: -d["rel_err"])[:3]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"],
                    help="storage type of the populations in HBM")
    ap.add_argument("--compute", default=None, choices=["float64", "float32"],
                    help="arithmetic type of the time-step kernel (default float64; float32 needs --dtype float32)")
    ap.add_argument("--api", default="auto", choices=["auto", "pylbm", "b200"],
                    help="front end: the reference's pylbm.Simulation + plugin (oracle/_ref), or pylbm_b200.Simulation")
    ap.add_argument("--in-place", action="store_true",
                    help="in-place streaming (AA pattern): ONE population array (single GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true", help="skip the N > 1 parity block")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo: direct NVLink peer stores from the fused kernel, or NCCL send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # the JSON line is the only thing on stdout: whatever the front ends, NCCL ("NCCL version ..." comes
    # from C) or the build tools print goes to stderr -- at the file-descriptor level
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    ctx = Context(args)
    rank, world = ctx.rank, ctx.world
    case_name, case_kw, description = WORKLOADS[args.workload]
    scaling = "weak" if args.workload in WEAK else "strong"
    if scaling == "weak":
        case_kw = dict(case_kw, nx=case_kw["nx"] * world)
    # identical in both arms (run-dependent facts go to "run")
    config = {"workload": description, "case": case_name, **case_kw, "storage": args.dtype,
              "arithmetic": args.compute or "float64", **({"streaming": "in place (one array)"} if args.in_place else {}),
              "l2": "working set (F + Fnew) far larger than the 126 MB L2; no flush needed"}

    # ---- reference arm: the reference's CPU implementation alone, rank 0 --------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        res = reference_cython_run((case_name, case_kw), args.steps, args.warmup)
        port = None
        if res is None:          # no oracle/_ref on this box: the restatement stands in
            res = cpu_port_run((case_name, case_kw), args.steps, args.warmup,
                               budget_s=float(os.environ.get("PYLBM_B200_CPU_BUDGET_S", "90")))
        elif not args.no_cpu_baseline:
            port = cpu_port_run((case_name, case_kw), 10, 2, budget_s=20.0)
        line = {
            "impl": "reference", "metric": "MLUPS", "value": res["value"], "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "cpu_port": {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")} if port else None,
            "e2e": {"value": res["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "run": {"api": "pylbm.Simulation generator='cython'" if res["kind"] == "reference" else "oracle port",
                    "parallelism": "1 rank x %d core(s)" % res["cores"], "halo": None, "setup_s": None},
        }
        print(json.dumps(line), file=out, flush=True)
        return 0

    import numpy as np

    ctx.init()
    lib, rt, dist = ctx.lib, ctx.rt, ctx.dist

    # ---- N > 1: parity of the slab path before anything is timed -----------------------------------
    parity = None
    if world > 1 and not args.no_parity:
        parity = multi_gpu_parity(ctx)

    t_build = time.time()
    sim = ctx.simulation(case_name, case_kw, args.dtype, args.compute)
    t_build = time.time() - t_build
    global_cells = float(np.prod(sim.domain.global_size))
    Q = sim.container.nv

    # ---- warm-up ----------------------------------------------------------
    sim.run(args.warmup)
    sim.synchronize()

    # ---- timed region: K steps, state resident in HBM ---------------------
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    elapsed_ms, launches = timed_run(ctx, sim, args.steps)
    samples_timed = len(sampler.lines) if rank == 0 else 0
    if elapsed_ms < 1500.0:
        # the timed region is too short for nvidia-smi to sample it: keep sampling while the SAME
        # load runs (untimed) so that the median clock under load is meaningful
        extra = int(max(1, min(20000, 1500.0 / max(elapsed_ms / args.steps, 1e-3))))
        sim.run(extra)
        sim.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["samples_in_timed_region"] = samples_timed
    value = global_cells * args.steps / (elapsed_ms * 1e-3) / 1e6

    # the same K steps as K calls of sol.one_time_step() (what a pylbm script does)
    stepwise_ms, _ = timed_run(ctx, sim, args.steps, stepwise=True)

    # ---- roofline of the fused kernel: per-launch CUDA events ----------------
    roofline = fused_kernel_roofline(ctx, sim, args.steps, elapsed_ms / args.steps, args.workload, args.dtype,
                                     args.compute)

    # ---- end to end through the public API with host buffers ------------------
    # every rank: its slab of F from pinned host memory -> HBM, K x sol.one_time_step(), its slab of
    # the conserved moments -> host; wall clock between two barriers, max over ranks
    e2e = None
    if not args.no_e2e and not args.in_place:
        F = sim.container.F
        nbytes = F.nv * int(np.prod(F.nspace)) * 8
        ptr = ctypes.c_void_p()
        rt.check(lib.lbm_host_alloc(ctypes.byref(ptr), nbytes), "lbm_host_alloc")
        host = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(F.nv,) + F.nspace)
        for k in range(F.nv):            # current state -> pinned host (outside the timed region)
            host[k] = F.get(k, 1)[0]
        sim.synchronize()
        ctx.barrier()
        k_e2e = args.steps
        t0 = time.perf_counter()
        sim.F_halo[slice(None)] = host   # H2D of the populations, pinned source (public item property)
        t_h2d = time.perf_counter() - t0
        for _ in range(k_e2e):
            sim.one_time_step()          # the call a user makes
        d2h = 0
        t1 = time.perf_counter()
        for key in sim.scheme.consm:     # conserved moments back on the host
            d2h += sim.m[key].nbytes
        sim.synchronize()
        t_d2h = time.perf_counter() - t1
        ctx.barrier()
        wall = time.perf_counter() - t0
        (wall,) = ctx.max_over_ranks(wall)
        h2d_total, d2h_total = ctx.sum_over_ranks(float(nbytes), float(d2h))
        e2e = {
            "value": global_cells * k_e2e / wall / 1e6, "unit": "MLUPS",
            "h2d_bytes_per_step": h2d_total / k_e2e, "d2h_bytes_per_step": d2h_total / k_e2e,
            "h2d_gbs_rank0": nbytes / t_h2d / 1e9, "wall_s": wall, "h2d_s_rank0": t_h2d, "d2h_and_steps_tail_s_rank0": t_d2h,
            "note": "timed (wall clock, max over ranks): pinned-host F -> HBM (%d bytes once, all ranks), %d x "
                    "sol.one_time_step(), conserved moments -> host (%d bytes once); the LBM state is resident "
                    "between steps, so the copies are amortised over the steps" % (h2d_total, k_e2e, d2h_total),
        }
        lib.lbm_host_free(ptr)
    del sim

    # ---- the other BASELINE configs, short ------------------------------------------------------------
    also = None
    if not args.no_also and args.workload == DEFAULT_WORKLOAD and args.dtype == "float64":
        also = []
        for workload, steps in (ALSO_1GPU if world == 1 else ALSO_NGPU):
            try:
                also.append(measure_also(ctx, workload, steps))
            except Exception as exc:            # the headline must survive a failing side measurement
                also.append({"workload": workload, "error": "%s: %s" % (type(exc).__name__, exc)})

    # ---- CPU baseline beside it (rank 0, N = 1) ---------------------------------
    cpu = cpu_port = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        try:
            res = reference_cython_run((case_name, case_kw), 10, 3)
        except Exception as exc:
            res = None
            cpu_port = {"reference_error": "%s: %s" % (type(exc).__name__, exc)}
        port = cpu_port_run((case_name, case_kw), 10, 2, budget_s=15.0)
        port = {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if res is not None:
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu_port = port
        else:
            cpu = port

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": ("f64" if args.dtype == "float64" else
                                                                    ("f32" if args.compute == "float32" else "f32-storage/f64-math")),
            "data": "synthetic", "config": config,
            "run": {"api": ctx.api, "parallelism": "x-slabs x%d" % world,
                    "halo": (args.halo if world > 1 else "periodic (single GPU)"), "setup_s": round(t_build, 2)},
            "stepwise": {"value": global_cells * args.steps / (stepwise_ms * 1e-3) / 1e6,
                         "ms_per_step": stepwise_ms / args.steps,
                         "note": "K Python calls of sol.one_time_step() instead of one sol.run(K)"},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "cpu_port": cpu_port, "gpu_launches": int(launches),
            "clocks": clocks, "frac_of_roofline": value / (world * roofline["roofline_mlups_per_gpu"]),
            "parity": parity, "also": also,
        }
        print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and rank == 0 and not parity["ok"]:
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
