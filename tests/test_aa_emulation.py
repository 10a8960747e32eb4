"""
In-place streaming (AA pattern) -- soundness by brute force on the CPU, before any CUDA.

ONE population array.  A time step pair is
  * even step ("gather / scatter back"): cell x reads A(k, x - v_k) like the pull kernel, collides, and writes
    its new population k into the slot it read for the OPPOSITE population: A(kbar, x + v_k).  The Q slots
    a cell touches belong to no other cell, so the step is in place.  A population that leaves the
    interior is stored twice: in the ghost slot (the boundary kernels of the next step read it there) and
    at the fully wrapped position (the periodic image, reference: storage.py:333-367 run backwards);
  * odd step ("local"): cell x reads A(kbar, x) -- that IS the value a pull would bring -- collides and
    writes A(k, x): the array is in the natural layout again (and the kernel stores the periodic images of
    the next even step like the two-array kernel does).
Boundary methods of an odd step run on transformed positions (k, y) -> (kbar, y + v_k), a bijection of the
slots, so their sequential semantics (levels, snapshots) carry over unchanged.

The emulator of tests/test_walls_emulation.py runs the reference order (two arrays) and this scheme on
random configurations built with the real front end; whenever `boundary.plan_aa` accepts, the interior
populations must be IDENTICAL after an even and after an odd number of steps.
"""
import numpy as np
import pytest

import test_walls_emulation as emu


def _setup(dico):
    from pylbm_b200.boundary import plan_aa

    dom, lay, vel, sym, methods, _ = emu._setup(dico)
    return dom, lay, vel, sym, methods, plan_aa(methods, lay, vel, sym)


def _run_aa(lay, vel, sym, methods, odd_methods, nsteps, f0):
    q = len(vel)
    n, w = lay.canonical_n, lay.canonical_vmax
    vel3 = np.zeros((q, 3), dtype=int)
    vel3[:, 3 - lay.dim:] = vel[:, : lay.dim]
    inner = tuple(slice(w[a], n[a] - w[a]) for a in range(3))
    idx = np.meshgrid(*[np.arange(w[a], n[a] - w[a]) for a in range(3)], indexing="ij")
    A = f0.copy()
    natural = True
    for _ in range(nsteps):
        if natural:
            emu._periodic(A, w, (0, 1, 2))
            for m in methods:
                emu._apply(A, m)
            pulled = [A[(k,) + tuple(slice(w[a] - vel3[k][a], n[a] - w[a] - vel3[k][a]) for a in range(3))].copy()
                      for k in range(q)]
            new = _collide(pulled)
            written = np.zeros(A.shape, dtype=np.int32)
            for k in range(q):
                tgt = [idx[a] + vel3[k][a] for a in range(3)]
                A[(sym[k],) + tuple(tgt)] = new[k]
                np.add.at(written, (np.full(tgt[0].shape, sym[k]),) + tuple(tgt), 1)
                out = np.zeros(tgt[0].shape, dtype=bool)
                wrapped = []
                for a in range(3):
                    nin = n[a] - 2 * w[a]
                    o = (tgt[a] < w[a]) | (tgt[a] >= n[a] - w[a])
                    out |= o
                    wrapped.append(np.where(o, (tgt[a] - w[a]) % max(nin, 1) + w[a], tgt[a]) if w[a] > 0 else tgt[a])
                if out.any():
                    sel = tuple(t[out] for t in wrapped)
                    A[(sym[k],) + sel] = new[k][out]
                    np.add.at(written, (np.full(sel[0].shape, sym[k]),) + sel, 1)
            assert written.max() <= 1, "two cells wrote the same slot"
            natural = False
        else:
            for m in odd_methods:
                emu._apply(A, m)
            pulled = [A[(sym[k],) + inner].copy() for k in range(q)]
            new = _collide(pulled)
            for k in range(q):
                A[(k,) + inner] = new[k]
            natural = True
    if not natural:     # natural view of the swapped array: S(k, x) = A(kbar, x + v_k)
        S = np.zeros_like(A)
        for k in range(q):
            S[(k,) + inner] = A[(sym[k],) + tuple(idx[a] + vel3[k][a] for a in range(3))]
        return S
    return A


def _collide(pulled):
    q = len(pulled)
    total = sum(pulled)
    return [0.75 * pulled[k] + 0.25 * total / q + 0.01 * pulled[k] * pulled[(k + 1) % q] for k in range(q)]


def _reference(dom, lay, vel, sym, methods, nsteps, f0):
    return emu._run(dom, lay, vel, sym, methods, None, nsteps, f0)


def _check(dico, rng, must_accept=None):
    dom, lay, vel, sym, methods, plan = _setup(dico)
    if must_accept is not None:
        assert (plan is not None) == must_accept
    if plan is None:
        return False
    q = len(vel)
    f0 = 1.0 / q + 0.05 * rng.uniform(-1, 1, size=(q,) + tuple(lay.canonical_n))
    w, n = lay.canonical_vmax, lay.canonical_n
    inner = (slice(None),) + tuple(slice(w[i], n[i] - w[i]) for i in range(3))
    for nsteps in (4, 5):
        a = _reference(dom, lay, vel, sym, methods, nsteps, f0)
        b = _run_aa(lay, vel, sym, methods, plan, nsteps, f0)
        assert np.array_equal(a[inner], b[inner]), nsteps
    return True


@pytest.mark.parametrize("dim,seed", [(2, s) for s in range(60)] + [(3, s) for s in range(25)] + [(1, s) for s in range(5)])
def test_in_place_streaming_reproduces_the_reference_order(dim, seed):
    rng = np.random.default_rng(1000 * dim + seed)
    dico = emu._random_case(rng, dim)
    if not _check(dico, rng):
        pytest.skip("plan refused")


def test_fully_periodic_and_two_ghost_layers():
    import pylbm_b200 as lb

    rng = np.random.default_rng(3)
    # fully periodic D2Q9 and D3Q19, and D2Q13 (|v| = 2: two ghost layers) with walls
    assert _check({"box": {"x": [0, 1], "y": [0, 0.75], "label": -1}, "space_step": 1 / 8,
                   "schemes": [{"velocities": list(range(9))}]}, rng, True)
    assert _check({"box": {"x": [0, 1], "y": [0, 0.75], "z": [0, 0.5], "label": -1}, "space_step": 1 / 8,
                   "schemes": [{"velocities": list(range(19))}]}, rng, True)
    bb = lb.bc.BounceBack
    assert _check({"box": {"x": [0, 1], "y": [0, 0.75], "label": [0, 0, -1, -1]}, "space_step": 1 / 8,
                   "schemes": [{"velocities": list(range(13))}],
                   "boundary_conditions": {0: {"method": {0: bb}}}}, rng, True)


@pytest.mark.parametrize("name", ["cavity2d_bb", "periodic_x_walls_y", "channel2d_inlet_outlet_obstacle",
                                  "channel2d_outlet_before_walls", "cavity3d_bb", "channel3d_d3q27",
                                  "bouzidi_walls", "neumann_top"])
def test_directed_in_place_configurations(name):
    dico, _ = emu._directed_cases()[name]
    rng = np.random.default_rng(5)
    assert _check(dico, rng), "every directed configuration is expected to run in place"


# ---------------------------------------------------------------------------
# in place + the fused walls of the fastest axis (boundary.plan_walls)
# ---------------------------------------------------------------------------
def _run_aa_walls(lay, vel, sym, methods, odd_methods, walls, masks, nsteps, f0):
    """the in-place scheme with the wall plan: the replaced entries run only on stale ghosts (first step);
    the even kernel stores `+-f_k(c) + rhs[k]` into the cell's OWN slot (k, c), the odd kernel into the
    wall's ghost cell (sym k, c + v_k), and neither keeps a periodic image along the fastest axis."""
    q = len(vel)
    n, w = lay.canonical_n, lay.canonical_vmax
    vel3 = np.zeros((q, 3), dtype=int)
    vel3[:, 3 - lay.dim:] = vel[:, : lay.dim]
    inner = tuple(slice(w[a], n[a] - w[a]) for a in range(3))
    idx = np.meshgrid(*[np.arange(w[a], n[a] - w[a]) for a in range(3)], indexing="ij")
    i0, i1 = np.meshgrid(np.arange(w[0], n[0] - w[0]), np.arange(w[1], n[1] - w[1]), indexing="ij")
    A = f0.copy()
    natural, fresh = True, False
    for _ in range(nsteps):
        if natural:
            if fresh:
                emu._periodic(A, w, (0, 1), keep_fast_ghosts=True)
            else:
                emu._periodic(A, w, (0, 1, 2))
            for im, m in enumerate(methods):
                emu._apply(A, m, ~masks[im])
                if not fresh:
                    emu._apply(A, m, masks[im])
            pulled = [A[(k,) + tuple(slice(w[a] - vel3[k][a], n[a] - w[a] - vel3[k][a]) for a in range(3))].copy()
                      for k in range(q)]
            new = _collide(pulled)
            for k in range(q):
                tgt = [idx[a] + vel3[k][a] for a in range(3)]
                A[(sym[k],) + tuple(tgt)] = new[k]
                out = np.zeros(tgt[0].shape, dtype=bool)
                wrapped = []
                for a in range(3):
                    nin = n[a] - 2 * w[a]
                    o = ((tgt[a] < w[a]) | (tgt[a] >= n[a] - w[a])) if (a != 2 and w[a] > 0) else np.zeros(tgt[a].shape, bool)
                    out |= o
                    wrapped.append(np.where(o, (tgt[a] - w[a]) % max(nin, 1) + w[a], tgt[a]))
                if out.any():
                    A[(sym[k],) + tuple(t[out] for t in wrapped)] = new[k][out]
            # walls, even step: the bounced value into the cell's own slot of population k
            for k in range(q):
                if vel3[k][2] == 0:
                    continue
                plane = walls["lo_plane"] if vel3[k][2] < 0 else walls["hi_plane"]
                neg = walls["neg_lo"] if vel3[k][2] < 0 else walls["neg_hi"]
                val = new[k][:, :, plane - w[2]]
                A[k, i0, i1, plane] = (-val if neg else val) + walls["rhs"][k]
            natural = False
        else:
            for im, m in enumerate(odd_methods):
                emu._apply(A, m, ~masks[im])
            pulled = [A[(sym[k],) + inner].copy() for k in range(q)]
            new = _collide(pulled)
            for k in range(q):
                A[(k,) + inner] = new[k]
            # walls, odd step: like the two-array walls kernel
            for k in range(q):
                if vel3[k][2] == 0:
                    continue
                plane = walls["lo_plane"] if vel3[k][2] < 0 else walls["hi_plane"]
                neg = walls["neg_lo"] if vel3[k][2] < 0 else walls["neg_hi"]
                val = new[k][:, :, plane - w[2]]
                A[sym[k], i0 + vel3[k][0], i1 + vel3[k][1], plane + vel3[k][2]] = (-val if neg else val) + walls["rhs"][k]
            natural, fresh = True, True
    if not natural:
        S = np.zeros_like(A)
        for k in range(q):
            S[(k,) + inner] = A[(sym[k],) + tuple(idx[a] + vel3[k][a] for a in range(3))]
        return S
    return A


@pytest.mark.parametrize("dim,seed", [(2, s) for s in range(60)] + [(3, s) for s in range(25)])
def test_in_place_streaming_with_the_wall_plan(dim, seed):
    from pylbm_b200.boundary import plan_aa

    rng = np.random.default_rng(1000 * dim + seed)
    dico = emu._random_case(rng, dim)
    if not dico["boundary_conditions"]:
        pytest.skip("fully periodic box")
    dom, lay, vel, sym, methods, plan = emu._setup(dico)
    if plan is None:
        pytest.skip("wall plan refused")
    walls, masks = plan
    odd = plan_aa(methods, lay, vel, sym)
    assert odd is not None
    q = len(vel)
    f0 = 1.0 / q + 0.05 * rng.uniform(-1, 1, size=(q,) + tuple(lay.canonical_n))
    w, n = lay.canonical_vmax, lay.canonical_n
    inner = (slice(None),) + tuple(slice(w[i], n[i] - w[i]) for i in range(3))
    for nsteps in (4, 5):
        a = emu._run(dom, lay, vel, sym, methods, None, nsteps, f0)
        b = _run_aa_walls(lay, vel, sym, methods, odd, walls, masks, nsteps, f0)
        assert np.array_equal(a[inner], b[inner]), nsteps
