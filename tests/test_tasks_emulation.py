"""
Soundness of the task plan (boundary.plan_tasks) by brute force, on the CPU.

Same emulator and random configurations as tests/test_walls_emulation.py (boundary lists from the real
front end: random face methods among bounce-back / anti-bounce-back / both Bouzidi kinds / Neumann,
periodic faces, an obstacle).  Two time loops:

* reference order (simulation.py:373-420): periodic update, the methods one after the other
  (sequential loops, Bouzidi bounce-back on a snapshot), pull + collision;
* what the fused kernel does with a task table: NO boundary loop; every (cell, population) named by a
  task pulls the value computed from the input array by the task's (kind, loads, rhs, coefficient),
  the block / thread / population packed in `code` and `block_ptr` are decoded the way the kernel does.

Whenever plan_tasks accepts, the interior populations must be IDENTICAL after several steps.
"""
import numpy as np
import pytest

import test_walls_emulation as emu


def _setup(dico):
    from pylbm_b200.boundary import plan_tasks

    dom, lay, vel, sym, methods, _ = emu._setup(dico)
    for m in methods:
        m["single"] = m["eligible"]
    return dom, lay, vel, methods, plan_tasks(methods, lay, vel)


def _run_tasks(lay, vel, methods, plan, nsteps, f0):
    q = len(vel)
    n, w = lay.canonical_n, lay.canonical_vmax
    vel3 = np.zeros((q, 3), dtype=int)
    vel3[:, 3 - lay.dim:] = vel[:, : lay.dim]
    tx, ngx, ngy = plan["tx"], plan["ngroups_x"], plan["ngroups_y"]
    ty = 128 // tx
    # decode (block, thread) -> cell like the kernel: block = ((i0 - w0) * ngy + by) * ngx + bx
    block = np.repeat(np.arange(plan["nblocks"]), np.diff(plan["block_ptr"]))
    code = plan["code"].astype(np.int64)
    thread, k, kind = code & 255, (code >> 8) & 255, code >> 16
    bx, rest = block % ngx, block // ngx
    by, bz = rest % ngy, rest // ngy
    i2 = w[2] + bx * tx + (thread & (tx - 1))
    i1 = w[1] + by * ty + thread // tx
    i0 = w[0] + bz
    assert (i2 < n[2] - w[2]).all() and (i1 < n[1] - w[1]).all() and (i0 < n[0] - w[0]).all()
    rhs = np.array([methods[i]["rhs"][j] for i, j in zip(plan["ibc"], plan["entry"])])
    f = f0.copy()
    for _ in range(nsteps):
        emu._periodic(f, w, (0, 1, 2))
        flat = f.reshape(-1)
        values = np.array([emu._bc_value(int(kd), flat[a], flat[b], r, d)
                           for kd, a, b, r, d in zip(kind, plan["l0"], plan["l1"], rhs, plan["dist"])])
        # the pull of population k at cell c reads (k, c - v_k): put the task values there
        g = f.copy()
        for t in range(len(values)):
            src = (int(k[t]), i0[t] - vel3[k[t]][0], i1[t] - vel3[k[t]][1], i2[t] - vel3[k[t]][2])
            g[src] = values[t]
        f = emu._pull_collide(g, vel3, w)
    return f


@pytest.mark.parametrize("dim,seed", [(2, s) for s in range(60)] + [(3, s) for s in range(25)] + [(1, s) for s in range(5)])
def test_accepted_task_plans_reproduce_the_reference_order(dim, seed):
    rng = np.random.default_rng(1000 * dim + seed)
    dico = emu._random_case(rng, dim)
    if not dico["boundary_conditions"]:
        pytest.skip("fully periodic box")
    dom, lay, vel, methods, plan = _setup(dico)
    if plan is None:
        pytest.skip("plan refused")
    assert 0 < plan["ntasks"] <= plan["nentries"]
    q = len(vel)
    f0 = 1.0 / q + 0.05 * rng.uniform(-1, 1, size=(q,) + tuple(lay.canonical_n))
    a = emu._run(dom, lay, vel, None, methods, None, 6, f0)
    b = _run_tasks(lay, vel, methods, plan, 6, f0)
    w, n = lay.canonical_vmax, lay.canonical_n
    inner = (slice(None),) + tuple(slice(w[i], n[i] - w[i]) for i in range(3))
    assert np.array_equal(a[inner], b[inner])


def test_the_fuzz_accepts_and_refuses_task_plans():
    seen = {(2, True): 0, (2, False): 0, (3, True): 0, (3, False): 0}
    for dim, count in ((2, 60), (3, 25)):
        for seed in range(count):
            dico = emu._random_case(np.random.default_rng(1000 * dim + seed), dim)
            if not dico["boundary_conditions"]:
                continue
            seen[(dim, _setup(dico)[-1] is not None)] += 1
    assert seen[(2, True)] >= 10 and seen[(3, True)] >= 5 and seen[(2, False)] + seen[(3, False)] >= 3, seen


@pytest.mark.parametrize("name,accepted", [
    ("cavity2d_bb", True), ("periodic_x_walls_y", True), ("channel2d_inlet_outlet_obstacle", True),
    ("channel2d_outlet_before_walls", False), ("cavity3d_bb", True), ("channel3d_d3q27", True),
    ("bouzidi_walls", True), ("neumann_top", True),
])
def test_directed_task_configurations(name, accepted):
    """plain and Bouzidi walls, an obstacle, periodic faces are independent entries; the edge entries of a
    Neumann outlet that copy wall values stored by an EARLIER method inherit those entries; an outlet
    applied BEFORE the walls reads what the walls overwrite later: refused (list kernels)."""
    dico, _ = emu._directed_cases()[name]
    dom, lay, vel, methods, plan = _setup(dico)
    assert (plan is not None) == accepted
    if plan is None:
        return
    rng = np.random.default_rng(11)
    q = len(vel)
    f0 = 1.0 / q + 0.05 * rng.uniform(-1, 1, size=(q,) + tuple(lay.canonical_n))
    a = emu._run(dom, lay, vel, None, methods, None, 5, f0)
    b = _run_tasks(lay, vel, methods, plan, 5, f0)
    w, n = lay.canonical_vmax, lay.canonical_n
    inner = (slice(None),) + tuple(slice(w[i], n[i] - w[i]) for i in range(3))
    assert np.array_equal(a[inner], b[inner])
