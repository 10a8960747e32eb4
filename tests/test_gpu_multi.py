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


def _demo_worker(rank, world, nccl_id, test, nsteps, queue, halo, shared, barrier, h5dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import pylbm_b200
    from pylbm_b200 import runtime as rt
    from demo_fixtures import load_demo

    rt.check(rt.lib().lbm_set_device(rank), "lbm_set_device")

    def gather(blob):
        shared[rank] = blob
        barrier.wait()
        out = [shared[r] for r in range(world)]
        barrier.wait()
        return out

    dico, kwargs, record = load_demo(test)
    sim = pylbm_b200.Simulation(dico, slab=(rank, world), nccl_id=nccl_id, gather=gather if halo == "peer" else None,
                                **kwargs)
    while sim.t < record["final_time"]:
        sim.one_time_step()
    assert sim.nt == nsteps
    out = {str(k): sim.m[k].copy() for k in sim.scheme.consm}
    if h5dir:
        # the slabs are brought to rank 0 by the topology's all-gather (reference: hdf5.py:163-178)
        h5 = pylbm_b200.H5File(sim.domain.mpi_topo, "slabs", h5dir)
        h5.set_grid(*sim.domain.coords)
        for k in sim.scheme.consm:
            h5.add_scalar(str(k), sim.m[k])
        h5.save()
    sim.synchronize()
    queue.put((rank, out))


@pytest.mark.parametrize("test,halo", [
    ("test2D_karman_vortex_street", "peer"), ("test2D_rayleigh_benard", "nccl"), ("test2D_orszag_Tang_vortex", "peer"),
    ("test3D_karman", "peer"), ("test3D_poseuille", "nccl"), ("test1D_euler", "peer"),
])
def test_reference_demos_on_slabs(test, halo, tmp_path):
    """dictionaries of the reference's demo tests, cut into x-slabs over the GPUs of the box, against the
    fields of the unmodified reference (fluid cells: solid cells next to an interface may differ)."""
    import ctypes
    import multiprocessing as mp

    import pylbm_b200
    from pylbm_b200 import runtime as rt
    from demo_fixtures import load_demo, load_results

    ngpu = rt.lib().lbm_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    expected = load_results(test)
    raw = (ctypes.c_char * 128)()
    rt.check(rt.lib().lbm_comm_unique_id(raw), "lbm_comm_unique_id")
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    manager = ctx.Manager()
    shared, barrier = manager.dict(), manager.Barrier(world)
    h5dir = str(tmp_path) if halo == "peer" else ""
    procs = [ctx.Process(target=_demo_worker, args=(r, world, bytes(raw.raw), test, expected["nsteps"], queue, halo,
                                                    shared, barrier, h5dir)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    dico, _, _ = load_demo(test)
    domain = pylbm_b200.Domain(dico)
    fluid = domain.in_or_out[tuple(slice(v, -v) for v in list(domain.stencil.vmax)[: domain.dim])] == domain.valin
    stride = expected["plane_stride"]
    tol = 1e-12 * max(1.0, expected["nsteps"] / 100.0)
    for key, want in expected["ref"].items():
        whole = np.concatenate([results[r][key] for r in range(world)], axis=0)
        mask = fluid
        if stride > 1:
            whole, mask = whole[::stride], fluid[::stride]
        assert whole.shape == want.shape
        err = np.abs(whole[mask] - want[mask]).max() / max(np.abs(want[mask]).max(), 1e-300)
        assert err <= tol, (test, key, err)
    if h5dir:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from make_golden import H5Lite

        back = H5Lite(os.path.join(h5dir, "slabs.h5")).datasets()
        for key in expected["ref"]:
            whole = np.concatenate([results[r][key] for r in range(world)], axis=0)
            assert np.array_equal(back[key], whole.T)


def _timeout_worker(rank, world, nccl_id, queue, shared, barrier):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ["PYLBM_B200_HALO_TIMEOUT_S"] = "2"
    import pylbm_b200
    from pylbm_b200 import cases, runtime as rt

    rt.check(rt.lib().lbm_set_device(rank), "lbm_set_device")

    def gather(blob):
        shared[rank] = blob
        barrier.wait()
        return [shared[r] for r in range(world)]

    sim = pylbm_b200.Simulation(cases.karman_d2q9(nx=64, ny=32), slab=(rank, world), nccl_id=nccl_id, gather=gather)
    sim.run(4)
    sim.synchronize()
    barrier.wait()
    message = ""
    if rank == 0:
        # rank 1 stops stepping (a dead rank): my next steps wait for a signal that never comes
        try:
            sim.run(3)
            sim.synchronize()
        except rt.LbmError as exc:
            message = str(exc)
    barrier.wait()
    queue.put((rank, message))


def test_a_silent_neighbour_is_reported_not_waited_for_ever():
    """k_wait gives up after PYLBM_B200_HALO_TIMEOUT_S seconds and the next runtime call says which
    neighbour was late (the default is 30 s); without it one dead rank hangs the node silently."""
    import ctypes
    import multiprocessing as mp

    from pylbm_b200 import runtime as rt

    if rt.lib().lbm_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    raw = (ctypes.c_char * 128)()
    rt.check(rt.lib().lbm_comm_unique_id(raw), "lbm_comm_unique_id")
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    manager = ctx.Manager()
    shared, barrier = manager.dict(), manager.Barrier(world)
    procs = [ctx.Process(target=_timeout_worker, args=(r, world, bytes(raw.raw), queue, shared, barrier))
             for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    assert "peer halo timeout" in results[0] and "neighbour" in results[0], results[0]


def _plugin_worker(rank, world, nccl_id, case, kw, nsteps, queue, shared, barrier):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from conftest import reference_paths

    for p in reversed(reference_paths()):
        sys.path.insert(0, p)
    import pylbm
    from pylbm_b200 import cases, plugin, runtime as rt

    rt.check(rt.lib().lbm_set_device(rank), "lbm_set_device")

    def gather(blob):
        shared[rank] = blob
        barrier.wait()
        out = [shared[r] for r in range(world)]
        barrier.wait()
        return out

    plugin.register()
    plugin.configure(slab=(rank, world), nccl_id=nccl_id, gather=gather)
    sol = pylbm.Simulation(cases.CASES[case](perturb=cases.WAVE, mod=pylbm, generator="cuda", **kw))
    for _ in range(nsteps):
        sol.one_time_step()
    out = {str(k): sol.m[k].copy() for k in sol.scheme.consm}
    sol.synchronize()
    queue.put((rank, out))


@pytest.mark.parametrize("case,kw", [("karman_d2q9", dict(nx=128, ny=32)), ("lid_cavity_d3q19", dict(n=16))])
def test_pylbm_simulation_on_slabs(case, kw):
    """the north-star interface, one process per GPU: `pylbm.Simulation(dico, generator='cuda')` cut into
    x-slabs (fused NVLink halo) reproduces the single-GPU run of the same interface."""
    import ctypes
    import multiprocessing as mp

    from conftest import reference_paths

    if reference_paths() is None:
        pytest.skip("reference pylbm not installed (tools/make_ref.sh)")
    for p in reversed(reference_paths()):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pylbm
    from pylbm_b200 import cases, plugin, runtime as rt

    if rt.lib().lbm_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world, nsteps = 2, 20
    raw = (ctypes.c_char * 128)()
    rt.check(rt.lib().lbm_comm_unique_id(raw), "lbm_comm_unique_id")
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    manager = ctx.Manager()
    shared, barrier = manager.dict(), manager.Barrier(world)
    procs = [ctx.Process(target=_plugin_worker, args=(r, world, bytes(raw.raw), case, kw, nsteps, queue, shared, barrier))
             for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    plugin.register()
    ref = pylbm.Simulation(cases.CASES[case](perturb=cases.WAVE, mod=pylbm, generator="cuda", **kw))
    for _ in range(nsteps):
        ref.one_time_step()
    fluid = ref.domain.in_or_out[tuple(slice(v, -v) for v in ref.domain.stencil.vmax)] == ref.domain.valin
    for key in ref.scheme.consm:
        whole = np.concatenate([results[r][str(key)] for r in range(world)], axis=0)
        full = ref.m[key]
        assert np.abs(whole[fluid] - full[fluid]).max() <= 1e-12 * np.abs(full[fluid]).max()


def _aa_worker(rank, world, nccl_id, case, kw, nsteps, queue):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import pylbm_b200
    from pylbm_b200 import cases, runtime as rt

    rt.check(rt.lib().lbm_set_device(rank), "lbm_set_device")
    sim = pylbm_b200.Simulation(cases.CASES[case](perturb=cases.WAVE, **kw), slab=(rank, world), nccl_id=nccl_id,
                                in_place=True)
    assert sim.container.Fnew is sim.container.F
    for _ in range(3):
        sim.one_time_step()
    sim.boundary_condition()
    sim.run(nsteps - 3)
    out = {"F%d" % k: sim.F[k].copy() for k in range(sim.container.nv)}
    out["swapped"] = sim._swapped
    sim.synchronize()
    queue.put((rank, out))


@pytest.mark.parametrize("case,kw,nsteps", [("karman_d2q9", dict(nx=128, ny=32), 12), ("lid_cavity_d3q19", dict(n=16), 9),
                                            ("shallow_water_d2q4", dict(n=32), 7)])
def test_in_place_streaming_on_slabs(case, kw, nsteps, monkeypatch):
    """in-place streaming (ONE array per GPU) on x-slabs with the NCCL halo: forward exchange before an even
    step, reverse exchange (ghost planes -> the neighbours' interior planes) after it; the protocol is
    proven in tests/test_aa_slabs_emulation.py.  Fluid populations IDENTICAL to the single-GPU two-array
    run of the same kernel library, after even and odd step counts."""
    import ctypes
    import multiprocessing as mp

    import pylbm_b200
    from pylbm_b200 import cases, runtime as rt

    ngpu = rt.lib().lbm_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    raw = (ctypes.c_char * 128)()
    rt.check(rt.lib().lbm_comm_unique_id(raw), "lbm_comm_unique_id")
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_aa_worker, args=(r, world, bytes(raw.raw), case, kw, nsteps, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    monkeypatch.setenv("PYLBM_B200_AA_LIBRARY", "1")
    ref = pylbm_b200.Simulation(cases.CASES[case](perturb=cases.WAVE, **kw))
    ref.run(nsteps)
    fluid = ref.domain.in_or_out[tuple(slice(v, -v) for v in ref.domain.stencil.vmax)] == ref.domain.valin
    assert results[0]["swapped"] == bool(nsteps % 2)
    for k in range(ref.container.nv):
        whole = np.concatenate([results[r]["F%d" % k] for r in range(world)], axis=0)
        assert np.array_equal(whole[fluid], ref.F[k][fluid]), "population %d" % k
