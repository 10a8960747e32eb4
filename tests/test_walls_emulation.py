"""
Soundness of the wall plan (boundary.plan_walls) by brute force, on the CPU.

A small emulator runs the time loop twice on random 2-D and 3-D configurations whose boundary lists come
from the real front-end (Domain + Boundary: random face labels with bounce-back / anti-bounce-back /
Neumann / Bouzidi methods, periodic faces, an obstacle, label-dependent right-hand sides):

* reference order (simulation.py:373-420): periodic update of every axis, the methods in order
  (sequential loops, Bouzidi bounce-back on a snapshot), pull + collision;
* what the runtime does when the plan is accepted: the fused kernel stores the bounced values of the
  cells next to the walls of the fastest axis together with its main stores and no longer produces the
  periodic images along that axis; only the remaining list entries run; the first step (ghost layers
  not produced by a fused launch) goes through the full lists, the replaced entries right after their
  owner method.

Whenever plan_walls accepts, the populations must be IDENTICAL after several steps.  The collision is
a cheap nonlinear mix of the pulled populations: only the data flow matters here.
"""
import numpy as np
import pytest


def _setup(dico):
    import pylbm_b200 as lb
    from pylbm_b200.boundary import Boundary, plan_walls, schedule
    from pylbm_b200.storage import Layout

    dom = lb.Domain(dico)
    stencil = dom.stencil
    bc = Boundary(dom, None, dico)
    nv = int(stencil.nv_ptr[-1])
    lay = Layout(nv, dom.shape_halo, list(dom.stencil.vmax), align=1)     # dense: position = flat index
    assert lay.lead == 0 and lay.pitch == lay.canonical_n[2]
    vel = np.asarray(stencil.get_all_velocities())
    sym = np.asarray(stencil.get_symmetric())
    methods = []
    for m in bc.methods:
        m.set_iload()
        m.fix_iload()
        store = lay.positions(m.istore.T)
        loads = [lay.positions(l.T) for l in m.iload]
        order, ptr, two = schedule(store, loads, snapshot=m.snapshot)
        # right-hand side: one value per (label, population), so that edge cells labelled by a
        # neighbouring face differ from the face they sit on
        k = m.istore[:, 0]
        rhs = 0.01 * (1 + np.asarray(m.ilabel)) * (1 + (k % 5)) if m.kind != 4 else np.zeros(len(k))
        dist = np.asarray(m.s) if hasattr(m, "s") else None
        methods.append({
            "kind": m.kind, "store": store[order], "loads": [l[order] for l in loads], "rhs": rhs[order],
            "dist": None if dist is None else dist[order], "snapshot": m.snapshot,
            "eligible": len(ptr) == 2 and not two[0],
        })
    plan = plan_walls(methods, lay, vel, sym)
    return dom, lay, vel, sym, methods, plan


def _bc_value(kind, a, b, rhs, d):
    if kind == 0:
        return a + rhs
    if kind == 1:
        return -a + rhs
    if kind == 2:
        return ((1.0 - d) * b + d * a) + rhs
    if kind == 3:
        return ((1.0 - d) * b - d * a) + rhs
    return a


def _apply(f, m, mask=None):
    """one method, sequential loop in list order (snapshot reads for Bouzidi bounce-back)."""
    flat = f.reshape(-1)
    src = flat.copy() if m["snapshot"] else flat
    idx = np.arange(len(m["store"])) if mask is None else np.nonzero(mask)[0]
    for i in idx:
        a = src[m["loads"][0][i]]
        b = src[m["loads"][1][i]] if len(m["loads"]) > 1 else 0.0
        d = m["dist"][i] if m["dist"] is not None else 0.0
        flat[m["store"][i]] = _bc_value(m["kind"], a, b, m["rhs"][i], d)


def _periodic(f, w, axes, keep_fast_ghosts=False):
    """ghost layers <- opposite interior layers, axis by axis over the full extent of the others."""
    for a in axes:
        if w[a] == 0:
            continue
        sl = [slice(None)] * 4
        if keep_fast_ghosts and a != 2:
            sl[3] = slice(w[2], f.shape[3] - w[2])      # images of x / y never touch the ghost rows of z
        n = f.shape[1 + a]
        lo, hi, src_hi, src_lo = list(sl), list(sl), list(sl), list(sl)
        lo[1 + a], src_hi[1 + a] = slice(0, w[a]), slice(n - 2 * w[a], n - w[a])
        hi[1 + a], src_lo[1 + a] = slice(n - w[a], n), slice(w[a], 2 * w[a])
        f[tuple(lo)] = f[tuple(src_hi)]
        f[tuple(hi)] = f[tuple(src_lo)]


def _pull_collide(f, vel3, w):
    q = f.shape[0]
    n = f.shape[1:]
    inner = tuple(slice(w[a], n[a] - w[a]) for a in range(3))
    pulled = []
    for k in range(q):
        sl = tuple(slice(w[a] - vel3[k][a], n[a] - w[a] - vel3[k][a]) for a in range(3))
        pulled.append(f[(k,) + sl])
    total = sum(pulled)
    fnew = np.zeros_like(f)
    for k in range(q):
        fnew[(k,) + inner] = 0.75 * pulled[k] + 0.25 * total / q + 0.01 * pulled[k] * pulled[(k + 1) % q]
    return fnew


def _run(dom, lay, vel, sym, methods, plan, nsteps, f0):
    q = len(vel)
    n = lay.canonical_n
    w = lay.canonical_vmax
    vel3 = np.zeros((q, 3), dtype=int)
    vel3[:, 3 - lay.dim:] = vel[:, : lay.dim]
    f = f0.copy()
    walls, masks = plan if plan is not None else (None, None)
    for step in range(nsteps):
        stale = step == 0 or walls is None
        if stale:
            _periodic(f, w, (0, 1, 2))
            for im, m in enumerate(methods):
                if walls is None:
                    _apply(f, m)
                else:                                   # remaining entries, then the replaced ones
                    _apply(f, m, ~masks[im])
                    _apply(f, m, masks[im])
        else:
            _periodic(f, w, (0, 1), keep_fast_ghosts=True)
            for im, m in enumerate(methods):
                _apply(f, m, ~masks[im])
        fnew = _pull_collide(f, vel3, w)
        if walls is not None:                           # the wall stores of the fused kernel
            flat = fnew.reshape(-1)
            i0, i1 = np.meshgrid(np.arange(w[0], n[0] - w[0]), np.arange(w[1], n[1] - w[1]), indexing="ij")
            rows = (i0 * n[1] + i1).ravel() * n[2]
            for k in range(q):
                if vel3[k][2] == 0:
                    continue
                plane = walls["lo_plane"] if vel3[k][2] < 0 else walls["hi_plane"]
                neg = walls["neg_lo"] if vel3[k][2] < 0 else walls["neg_hi"]
                cell = rows + plane
                voff = (vel3[k][0] * n[1] + vel3[k][1]) * n[2] + vel3[k][2]
                a = flat[k * lay.pstride + cell]
                flat[sym[k] * lay.pstride + cell + voff] = (-a if neg else a) + walls["rhs"][k]
        f = fnew
    return f


def _random_case(rng, dim):
    import pylbm_b200 as lb

    bc = lb.bc
    kinds = [bc.BounceBack, bc.AntiBounceBack, bc.BouzidiBounceBack, bc.BouzidiAntiBounceBack]
    neumann = [bc.NeumannX, bc.NeumannY, bc.NeumannZ]
    n = [int(rng.integers(5, 9)) for _ in range(dim)]
    dx = 1.0 / n[-1]
    box = {"x": [0.0, n[0] * dx], "label": []}
    if dim > 1:
        box["y"] = [0.0, n[1] * dx]
    if dim > 2:
        box["z"] = [0.0, n[2] * dx]
    labels, conditions = [None] * (2 * dim), {}
    # the fastest axis gets the smallest labels most of the time, so that its method precedes the others
    order = [dim - 1] + list(range(dim - 1)) if rng.random() < 0.7 else list(range(dim))
    for axis in order:
        if axis < dim - 1 and rng.random() < 0.2:
            labels[2 * axis: 2 * axis + 2] = [-1, -1]
            continue
        for side in range(2):
            lab = len(conditions)
            r = rng.random()
            if axis == dim - 1:                       # the fastest axis: mostly plain walls
                method = kinds[int(rng.integers(0, 2))] if r < 0.85 else kinds[int(rng.integers(0, 4))]
            elif r < 0.25:
                method = neumann[axis]
            else:
                method = kinds[int(rng.integers(0, 4))]
            conditions[lab] = {"method": {0: method}}
            labels[2 * axis + side] = lab
    if dim > 1 and rng.random() < 0.5 and labels[-2] is not None and labels[-2] >= 0:
        # same kind on both walls of the fastest axis more often (the plan needs one kind per face only)
        conditions[labels[-1]] = dict(conditions[labels[-2]])
    box["label"] = labels
    dico = {"box": box, "space_step": dx, "scheme_velocity": 1.0,
            "schemes": [{"velocities": list(range({1: 3, 2: 9, 3: 19}[dim]))}],
            "boundary_conditions": conditions}
    if dim == 2 and min(n) >= 7 and rng.random() < 0.4:
        lab = len(conditions)
        conditions[lab] = {"method": {0: bc.BouzidiBounceBack}}
        dico["elements"] = [lb.Circle([0.5 * n[0] * dx, 0.5 * n[1] * dx], 1.3 * dx, label=lab)]
    return dico


@pytest.mark.parametrize("dim,seed", [(2, s) for s in range(60)] + [(3, s) for s in range(25)] + [(1, s) for s in range(5)])
def test_accepted_plans_reproduce_the_reference_order(dim, seed):
    rng = np.random.default_rng(1000 * dim + seed)
    dico = _random_case(rng, dim)
    if not dico["boundary_conditions"]:
        pytest.skip("fully periodic box")
    dom, lay, vel, sym, methods, plan = _setup(dico)
    if plan is None:
        pytest.skip("plan refused")
    q = len(vel)
    f0 = 1.0 / q + 0.05 * rng.uniform(-1, 1, size=(q,) + tuple(lay.canonical_n))
    a = _run(dom, lay, vel, sym, methods, None, 6, f0)
    b = _run(dom, lay, vel, sym, methods, plan, 6, f0)
    w, n = lay.canonical_vmax, lay.canonical_n
    inner = (slice(None),) + tuple(slice(w[i], n[i] - w[i]) for i in range(3))
    assert np.array_equal(a[inner], b[inner])


def test_the_fuzz_accepts_and_refuses():
    """the random generator above must exercise both outcomes, in 2-D and in 3-D."""
    seen = {(2, True): 0, (2, False): 0, (3, True): 0, (3, False): 0}
    for dim, count in ((2, 60), (3, 25)):
        for seed in range(count):
            dico = _random_case(np.random.default_rng(1000 * dim + seed), dim)
            if not dico["boundary_conditions"]:
                continue
            plan = _setup(dico)[-1]
            seen[(dim, plan is not None)] += 1
    assert min(seen.values()) >= 3, seen


def _directed_cases():
    import pylbm_b200 as lb

    bc = lb.bc
    BB, ABB, BZ, NX = bc.BounceBack, bc.AntiBounceBack, bc.BouzidiBounceBack, bc.NeumannX

    def box2(nx, ny, labels):
        return {"x": [0.0, nx / ny], "y": [0.0, 1.0], "label": labels}

    def box3(nx, ny, nz, labels):
        return {"x": [0.0, nx / nz], "y": [0.0, ny / nz], "z": [0.0, 1.0], "label": labels}

    d2 = [{"velocities": list(range(9))}]
    d3 = [{"velocities": list(range(19))}]
    d3q27 = [{"velocities": list(range(27))}]
    return {
        "cavity2d_bb": ({"box": box2(7, 6, [0, 0, 0, 1]), "space_step": 1 / 6, "schemes": d2,
                         "boundary_conditions": {0: {"method": {0: BB}}, 1: {"method": {0: BB}}}}, True),
        # diagonal links through a corner shared with a PERIODIC face keep their periodic image in the
        # reference (the corner ghost cell carries no boundary entry): refused
        "periodic_x_walls_y": ({"box": box2(8, 6, [-1, -1, 0, 1]), "space_step": 1 / 6, "schemes": d2,
                                "boundary_conditions": {0: {"method": {0: ABB}}, 1: {"method": {0: BB}}}}, False),
        "periodic_x_walls_y_d2q5": ({"box": box2(8, 6, [-1, -1, 0, 1]), "space_step": 1 / 6,
                                     "schemes": [{"velocities": list(range(5))}],
                                     "boundary_conditions": {0: {"method": {0: ABB}}, 1: {"method": {0: BB}}}}, True),
        # walls first (label 0): the Neumann outlet, applied later, reads the wall values at the edge
        "channel2d_inlet_outlet_obstacle": (
            {"box": box2(12, 8, [1, 2, 0, 0]), "space_step": 1 / 8, "schemes": d2,
             "elements": [lb.Circle([0.5, 0.5], 0.16, label=3)],
             "boundary_conditions": {0: {"method": {0: BB}}, 1: {"method": {0: BB}}, 2: {"method": {0: NX}},
                                     3: {"method": {0: BZ}}}}, True),
        # the outlet method comes BEFORE the walls: in the reference its edge entries read the periodic
        # image, not the wall value -- refused
        "channel2d_outlet_before_walls": (
            {"box": box2(12, 8, [0, 1, 2, 2]), "space_step": 1 / 8, "schemes": d2,
             "boundary_conditions": {0: {"method": {0: BZ}}, 1: {"method": {0: NX}}, 2: {"method": {0: BB}}}}, False),
        "cavity3d_bb": ({"box": box3(6, 5, 6, [0, 0, 0, 0, 0, 1]), "space_step": 1 / 6, "schemes": d3,
                         "boundary_conditions": {0: {"method": {0: BB}}, 1: {"method": {0: BB}}}}, True),
        "channel3d_d3q27": ({"box": box3(8, 5, 6, [1, 2, 0, 0, 0, 0]), "space_step": 1 / 6, "schemes": d3q27,
                             "boundary_conditions": {0: {"method": {0: BB}}, 1: {"method": {0: BB}},
                                                     2: {"method": {0: NX}}}}, True),
        "bouzidi_walls": ({"box": box2(7, 6, [0, 0, 0, 0]), "space_step": 1 / 6, "schemes": d2,
                           "boundary_conditions": {0: {"method": {0: BZ}}}}, False),
        "neumann_top": ({"box": box2(7, 6, [0, 0, 0, 1]), "space_step": 1 / 6, "schemes": d2,
                         "boundary_conditions": {0: {"method": {0: BB}}, 1: {"method": {0: bc.NeumannY}}}}, False),
    }


@pytest.mark.parametrize("name", ["cavity2d_bb", "periodic_x_walls_y", "periodic_x_walls_y_d2q5",
                                  "channel2d_inlet_outlet_obstacle", "channel2d_outlet_before_walls",
                                  "cavity3d_bb", "channel3d_d3q27", "bouzidi_walls", "neumann_top"])
def test_directed_configurations(name):
    """typical set-ups: the plan must be accepted where the walls of the fastest axis are plain
    (anti-)bounce-back -- also next to an inlet / outlet / obstacle handled by other methods -- and
    refused for Bouzidi or Neumann walls; accepted plans must reproduce the reference order."""
    dico, accepted = _directed_cases()[name]
    dom, lay, vel, sym, methods, plan = _setup(dico)
    assert (plan is not None) == accepted
    if plan is None:
        return
    rng = np.random.default_rng(7)
    q = len(vel)
    f0 = 1.0 / q + 0.05 * rng.uniform(-1, 1, size=(q,) + tuple(lay.canonical_n))
    a = _run(dom, lay, vel, sym, methods, None, 6, f0)
    b = _run(dom, lay, vel, sym, methods, plan, 6, f0)
    w, n = lay.canonical_vmax, lay.canonical_n
    inner = (slice(None),) + tuple(slice(w[i], n[i] - w[i]) for i in range(3))
    assert np.array_equal(a[inner], b[inner])
    replaced = sum(int(m.sum()) for m in plan[1])
    assert replaced > 0
