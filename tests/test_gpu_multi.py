"""Slab decomposition over several GPUs of one box (NCCL send/recv): N x-slabs must reproduce the
single-GPU run.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, nccl_id, case, kw, nsteps, queue, halo, shared, barrier):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import pylbm_b200
    from pylbm_b200 import cases, runtime as rt

    rt.check(rt.lib().lbm_set_device(rank), "lbm_set_device")

    def gather(blob):
        shared[rank] = blob
        barrier.wait()
        return [shared[r] for r in range(world)]

    sim = pylbm_b200.Simulation(cases.CASES[case](perturb=cases.WAVE, **kw), slab=(rank, world), nccl_id=nccl_id,
                                gather=gather if halo == "peer" else None)
    sim.run(7)              # graph pairs + single steps
    sim.boundary_condition()   # consumes the neighbours' signal once; the next step must not wait again
    for _ in range(3):
        sim.one_time_step()
    sim.boundary_condition()
    sim.boundary_condition()
    sim.F_halo[0] = sim.F_halo[0]   # outside write: ghosts invalidated, next step exchanges again
    sim.run(nsteps - 10)
    out = {str(k): sim.m[k].copy() for k in sim.scheme.consm}
    sim.synchronize()
    queue.put((rank, out))


@pytest.mark.parametrize("halo", ["peer", "nccl"])
@pytest.mark.parametrize("case,kw", [("karman_d2q9", dict(nx=128, ny=32)), ("lid_cavity_d3q19", dict(n=16)),
                                     ("channel_sphere_d3q27", dict(nx=32, ny=16, nz=16))])
def test_slabs_reproduce_single_gpu(case, kw, halo):
    import ctypes
    import multiprocessing as mp

    import pylbm_b200
    from pylbm_b200 import cases, runtime as rt

    ngpu = rt.lib().lbm_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    nsteps = 30
    raw = (ctypes.c_char * 128)()
    rt.check(rt.lib().lbm_comm_unique_id(raw), "lbm_comm_unique_id")
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    manager = ctx.Manager()
    shared, barrier = manager.dict(), manager.Barrier(world)
    procs = [ctx.Process(target=_worker, args=(r, world, bytes(raw.raw), case, kw, nsteps, queue, halo, shared, barrier))
             for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    ref = pylbm_b200.Simulation(cases.CASES[case](perturb=cases.WAVE, **kw))
    ref.run(nsteps)
    fluid = ref.domain.in_or_out[tuple(slice(v, -v) for v in ref.domain.stencil.vmax)] == ref.domain.valin
    for key in ref.scheme.consm:
        whole = np.concatenate([results[r][str(key)] for r in range(world)], axis=0)
        full = ref.m[key]
        assert whole.shape == full.shape
        assert np.abs(whole[fluid] - full[fluid]).max() <= 1e-12 * np.abs(full[fluid]).max()
