"""
Multi-rank path on the CPU: two processes (gloo, world_size 2), each owning one x-slab, built with the
same host logic the GPU path uses (SlabTopology -> Domain with interface label -2 -> per-slab boundary
lists) and exchanging ghost planes with the protocol of the runtime's NCCL exchange (sign-matched
populations; planes [w, 2w) to the left neighbour's high ghost, [n-2w, n-w) to the right neighbour's
low ghost; periodic ring).  The per-slab oracle must reproduce the single-rank oracle on fluid cells.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, kw, nsteps, queue):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    from pylbm_b200 import cases
    from pylbm_b200.domain import SlabTopology
    from oracle.lbm_oracle import OracleSimulation

    dist.init_process_group("gloo", rank=rank, world_size=world)
    final_time = None
    if case.startswith("fixture:"):
        from demo_fixtures import load_demo

        dico, _, record = load_demo(case[len("fixture:"):])
        final_time = record["final_time"]
    else:
        dico = cases.CASES[case](perturb=cases.WAVE, **kw)
    topo = SlabTopology(len(dico["box"]) - 1, rank, world)
    left, right = topo.left, topo.right
    vel = None

    def exchange(f, w):
        # f: AoS [x, y, z, Q]; same selection and pairing as exchange_slabs() in csrc/lbm_runtime.cu
        n = f.shape[0]
        plus = [k for k in range(f.shape[-1]) if vel[k][0] > 0]
        minus = [k for k in range(f.shape[-1]) if vel[k][0] < 0]
        to_left = torch.from_numpy(np.ascontiguousarray(f[w:2 * w][..., minus]))
        to_right = torch.from_numpy(np.ascontiguousarray(f[n - 2 * w:n - w][..., plus]))
        from_right = torch.empty_like(to_left)
        from_left = torch.empty_like(to_right)
        reqs = [dist.irecv(from_right, src=right, tag=1), dist.irecv(from_left, src=left, tag=2),
                dist.isend(to_left, dst=left, tag=1), dist.isend(to_right, dst=right, tag=2)]
        for r in reqs:
            r.wait()
        f[n - w:][..., minus] = from_right.numpy()
        f[:w][..., plus] = from_left.numpy()

    sim = OracleSimulation(dico, topology=topo, exchange=exchange)
    vel = sim.scheme.stencil.get_all_velocities()
    if final_time is not None:
        while sim.t < final_time:          # the loop of the reference demos
            sim.one_time_step()
    else:
        for _ in range(nsteps):
            sim.one_time_step()
    out = {str(k): sim.m[k].copy() for k in sim.scheme.consm}
    out["nt"] = sim.nt
    out["region"] = sim.domain.region[0]
    out["labels"] = sim.domain.box_label
    out["ncond"] = [int(m.istore.shape[0]) for m in sim.bc.methods]
    queue.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,kw", [("karman_d2q9", dict(nx=64, ny=32)), ("lid_cavity_d3q19", dict(n=12))])
def test_two_slabs_reproduce_one_rank(case, kw):
    import torch.multiprocessing as mp

    from pylbm_b200 import cases
    from oracle.lbm_oracle import OracleSimulation

    nsteps, world = 20, 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, kw, nsteps, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    ref = OracleSimulation(cases.CASES[case](perturb=cases.WAVE, **kw))
    for _ in range(nsteps):
        ref.one_time_step()
    inner = tuple(slice(v, -v) for v in ref.domain.stencil.vmax)
    fluid = ref.domain.in_or_out[inner] == ref.domain.valin
    # interface faces carry label -2 and no boundary entries (reference: domain.py:277-282, boundary.py:84,106)
    assert results[0]["labels"][1] == -2 and results[1]["labels"][0] == -2
    assert sum(sum(r["ncond"]) for r in results.values()) == sum(int(m.istore.shape[0]) for m in ref.bc.methods)
    for key in ref.scheme.consm:
        whole = np.concatenate([results[r][str(key)] for r in range(world)], axis=0)
        full = ref.m[key]
        assert whole.shape == full.shape
        scale = np.abs(full[fluid]).max()
        assert np.abs(whole[fluid] - full[fluid]).max() <= 1e-12 * max(scale, 1e-300)


@pytest.mark.parametrize("test", ["test2D_rayleigh_benard", "test3D_poseuille", "test1D_euler", "test2D_coude"])
def test_two_slabs_reproduce_the_reference_demo_fields(test):
    """dictionaries of the reference's demo tests on two gloo ranks (x-slabs, oracle kernels, the
    runtime's exchange protocol) against the fields of the unmodified single-rank reference:
    time-dependent boundary values, vectorial schemes, 1-D and obstacles cut by nothing but the slabs."""
    import torch.multiprocessing as mp

    import pylbm_b200
    from demo_fixtures import load_demo, load_results

    world = 2
    expected = load_results(test)
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, "fixture:" + test, {}, 0, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dico, _, _ = load_demo(test)
    domain = pylbm_b200.Domain(dico)
    fluid = domain.in_or_out[tuple(slice(v, -v) for v in list(domain.stencil.vmax)[: domain.dim])] == domain.valin
    for r in range(world):
        assert results[r]["nt"] == expected["nsteps"]
    for key, want in expected["ref"].items():
        whole = np.concatenate([results[r][key] for r in range(world)], axis=0)
        assert whole.shape == want.shape
        err = np.abs(whole[fluid] - want[fluid]).max() / max(np.abs(want[fluid]).max(), 1e-300)
        assert err <= 1e-12, (test, key, err)
