#!/usr/bin/env python
"""
bench.py -- MLUPS of the fused lattice Boltzmann time step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

One JSON line on rank 0.  A "step" is one `one_time_step` of the whole lattice:
ghost update -> boundary kernels -> fused pull stream+collide -> swap.

Default workload = BASELINE config 4: D3Q19 MRT lid-driven cavity 512^3, fp64, the
configuration the north-star target is quoted on; it fits one B200 (41 GB for F+Fnew).
For N > 1 the SAME global lattice is cut into N x-slabs (strong scaling), one process
per GPU, halo planes exchanged with NCCL send/recv.

  value        MLUPS = K * interior cells / device time (max over ranks), state resident in HBM
  roofline     fused kernel only: 2*Q*8 bytes per cell * cells / (CUDA-event time per launch)
               against MEASURED_PEAKS.json hbm_gbs
  e2e          same metric through the public API with HOST buffers: pinned host F -> device,
               K x sol.one_time_step(), conserved moments -> host (copies inside the timed region)
  cpu_baseline the oracle's C restatement of the reference Cython path on the host cores,
               bounded sample (rank 0, N = 1 only)

--impl reference times that CPU restatement alone (all host threads), same metric/config.
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


# ---------------------------------------------------------------------------
# CPU arm: the oracle's C restatement of the reference Cython path
# ---------------------------------------------------------------------------
def cpu_reference_run(case, steps, warmup, budget_s=60.0, threads=None):
    """MLUPS of the CPU restatement on a bounded sample of the workload."""
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
    rate = 3 * float(__import__("numpy").prod(sim.domain.shape_in)) / (time.perf_counter() - t0)  # cells/s
    # largest sample (not above the real workload) whose run fits the budget
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
    import numpy as np

    cells = float(np.prod(sim.domain.shape_in))
    return {
        "value": cells * steps / dt / 1e6,
        "unit": "MLUPS",
        "cores": threads,
        "kind": "port",
        "sample": "%s %s, %d steps after %d warm-up, %.1f s, OpenMP C restatement of the reference "
                  "Cython one_time_step (oracle/lbm_oracle.py), %s" % (name, kws, steps, warmup, dt, cpu_model()),
        "ms_per_step": dt / steps * 1e3,
    }


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo: direct NVLink peer stores from the fused kernel, or NCCL send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    case_name, case_kw, description = WORKLOADS[args.workload]
    scaling = "weak" if args.workload in WEAK else "strong"
    if scaling == "weak":
        case_kw = dict(case_kw, nx=case_kw["nx"] * world)
    q_of = {"lid_cavity_d3q19": 19, "karman_d2q9": 9, "shallow_water_d2q4": 12, "lid_cavity_d2q9": 9,
            "channel_sphere_d3q27": 27}
    config = {"workload": description, "case": case_name, **case_kw, "storage": args.dtype,
              "arithmetic": args.compute or "float64",
              "l2": "working set (F + Fnew) far larger than the 126 MB L2; no flush needed"}

    # ---- reference arm: CPU restatement only, rank 0 alone -----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        res = cpu_reference_run((case_name, case_kw), args.steps, args.warmup,
                                budget_s=float(os.environ.get("PYLBM_B200_CPU_BUDGET_S", "90")))
        line = {
            "impl": "reference", "metric": "MLUPS", "value": res["value"], "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import numpy as np
    import pylbm_b200
    from pylbm_b200 import cases, runtime as rt

    lib = rt.lib()
    dist = None
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        rt.check(lib.lbm_set_device(local_rank), "lbm_set_device")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (ctypes.c_char * 128)()
            rt.check(lib.lbm_comm_unique_id(raw), "lbm_comm_unique_id")
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        nccl_id = bytes(buf.cpu().numpy().tobytes())
    else:
        rt.check(lib.lbm_set_device(local_rank), "lbm_set_device")

    def barrier():
        if dist is not None:
            dist.barrier()

    dico = cases.CASES[case_name](**case_kw)
    t_build = time.time()
    gather = None
    if world > 1 and args.halo == "peer":
        def gather(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
    sim = pylbm_b200.Simulation(dico, dtype=args.dtype, slab=(rank, world) if world > 1 else None, nccl_id=nccl_id,
                                gather=gather, compute_dtype=args.compute)
    t_build = time.time() - t_build
    global_cells = float(np.prod(sim.domain.global_size))
    local_cells = float(np.prod(sim.domain.shape_in))
    Q = sim.container.nv
    itemsize = sim.container.F.itemsize

    # ---- warm-up ----------------------------------------------------------
    sim.run(args.warmup)
    sim.synchronize()

    # ---- timed region: K steps, state resident in HBM ---------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    sim.synchronize()
    launches0 = lib.lbm_sim_launch_count(sim._handle)
    rt.check(lib.lbm_sim_timer_start(sim._handle), "timer_start")
    sim.run(args.steps)
    ms = ctypes.c_float()
    rt.check(lib.lbm_sim_timer_stop(sim._handle, ctypes.byref(ms)), "timer_stop")
    sim.synchronize()
    barrier()
    launches = lib.lbm_sim_launch_count(sim._handle) - launches0
    elapsed_ms = float(ms.value)
    if dist is not None:
        import torch

        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
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

    # ---- roofline of the fused kernel: per-launch CUDA events ----------------
    rt.check(lib.lbm_sim_profile(sim._handle, 1), "profile")
    sim.run(args.steps, graph=False)
    fused_ms, nl = ctypes.c_double(), ctypes.c_int64()
    rt.check(lib.lbm_sim_profile_read(sim._handle, ctypes.byref(fused_ms), ctypes.byref(nl)), "profile_read")
    rt.check(lib.lbm_sim_profile(sim._handle, 0), "profile")
    per_launch_ms = fused_ms.value / max(nl.value, 1)
    bytes_per_launch = 2.0 * Q * itemsize * local_cells
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic = None
    variant = "" if args.dtype == "float64" else ("_f32" if args.compute == "float32" else "_f32storage")
    tpath = os.path.join(ROOT, "profiles", "traffic_%s%s.json" % (args.workload, variant))
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "kernel": "lbmk_kernel_one_time_step", "launch_ms": per_launch_ms,
        "kernel_share_of_step": per_launch_ms / (elapsed_ms / args.steps) if world == 1 else None,
        "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
        "roofline_mlups_per_gpu": peak * 1e9 / (2.0 * Q * itemsize) / 1e6,
    }

    # ---- end to end through the public API with host buffers ------------------
    # every rank: its slab of F from pinned host memory -> HBM, K x sol.one_time_step(), its slab of
    # the conserved moments -> host; wall clock between two barriers, max over ranks
    e2e = None
    if not args.no_e2e:
        F = sim.container.F
        nbytes = F.nv * int(np.prod(F.nspace)) * 8
        ptr = ctypes.c_void_p()
        rt.check(lib.lbm_host_alloc(ctypes.byref(ptr), nbytes), "lbm_host_alloc")
        host = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(F.nv,) + F.nspace)
        for k in range(F.nv):            # current state -> pinned host (outside the timed region)
            host[k] = F.get(k, 1)[0]
        sim.synchronize()
        barrier()
        k_e2e = args.steps
        t0 = time.perf_counter()
        sim.container.F.set(host)        # H2D of the populations, pinned source
        sim.container.Fnew.copy_from(sim.container.F)
        for _ in range(k_e2e):
            sim.one_time_step()          # the call a user makes
        d2h = 0
        for key in sim.scheme.consm:     # conserved moments back on the host
            d2h += sim.m[key].nbytes
        sim.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        h2d_total, d2h_total = float(nbytes), float(d2h)
        if dist is not None:
            import torch

            t = torch.tensor([wall, h2d_total, d2h_total], device="cuda", dtype=torch.float64)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            wall, h2d_total, d2h_total = float(tmax[0].item()), float(t[1].item()), float(t[2].item())
        e2e = {
            "value": global_cells * k_e2e / wall / 1e6, "unit": "MLUPS",
            "h2d_bytes_per_step": h2d_total / k_e2e, "d2h_bytes_per_step": d2h_total / k_e2e,
            "note": "timed (wall clock, max over ranks): pinned-host F -> HBM (%d bytes once, all ranks), %d x "
                    "sol.one_time_step(), conserved moments -> host (%d bytes once); the LBM state is resident "
                    "between steps, so the copies are amortised over the steps" % (h2d_total, k_e2e, d2h_total),
        }
        lib.lbm_host_free(ptr)

    # ---- CPU baseline beside it (rank 0, N = 1) ---------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        res = cpu_reference_run((case_name, case_kw), 10, 2, budget_s=20.0)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": ("f64" if args.dtype == "float64" else
                                                                    ("f32" if args.compute == "float32" else "f32-storage/f64-math")),
            "data": "synthetic", "config": dict(config, parallelism="x-slabs x%d" % world, setup_s=round(t_build, 2),
                           halo=(args.halo if world > 1 else "periodic (single GPU)")),
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches),
            "clocks": clocks, "frac_of_roofline": value / (world * roofline["roofline_mlups_per_gpu"]),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
