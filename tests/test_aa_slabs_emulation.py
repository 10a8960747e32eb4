"""
In-place streaming (AA pattern) on x-slabs -- soundness of the exchange protocol by brute force on the
CPU (NumPy emulator of tests/test_aa_emulation.py, two slabs in one process, boundary lists of every
slab from the real front end with SlabTopology: interface faces carry label -2).

Per pair of steps and per slab:
  even step:  FORWARD exchange (interior planes [w, 2w) / [n-2w, n-w) -> the neighbours' ghost planes,
              sign-matched populations, as in the two-array scheme), periodic update of the other axes,
              boundary methods, gather + scatter back (no wrap along the slab axis);
              REVERSE exchange: the ghost planes -- where the even step deposited the populations that
              left through the slab faces -- travel to the neighbours' interior planes [n-2w, n-w) / [w, 2w),
              the same sign-matched populations;
  odd step:   boundary methods on the transformed lists, local kernel.
Interior populations must be IDENTICAL to the single-domain two-array reference order.
"""
import numpy as np
import pytest

import test_walls_emulation as emu
from test_aa_emulation import _collide


def _slab_setup(dico, rank, size):
    import pylbm_b200 as lb
    from pylbm_b200.boundary import Boundary, plan_aa, schedule
    from pylbm_b200.domain import SlabTopology
    from pylbm_b200.storage import Layout

    dim = len(dico["box"]) - 1
    dom = lb.Domain(dico, topology=SlabTopology(dim, rank, size))
    stencil = dom.stencil
    bc = Boundary(dom, None, dico)
    nv = int(stencil.nv_ptr[-1])
    lay = Layout(nv, dom.shape_halo, list(stencil.vmax), align=1)
    vel = np.asarray(stencil.get_all_velocities())
    sym = np.asarray(stencil.get_symmetric())
    methods = []
    for m in bc.methods:
        m.set_iload()
        m.fix_iload()
        store = lay.positions(m.istore.T)
        loads = [lay.positions(l.T) for l in m.iload]
        order, ptr, two = schedule(store, loads, snapshot=m.snapshot)
        k = m.istore[:, 0]
        rhs = 0.01 * (1 + np.asarray(m.ilabel)) * (1 + (k % 5)) if m.kind != 4 else np.zeros(len(k))
        dist = np.asarray(m.s) if hasattr(m, "s") else None
        methods.append({"kind": m.kind, "store": store[order], "loads": [l[order] for l in loads], "rhs": rhs[order],
                        "dist": None if dist is None else dist[order], "snapshot": m.snapshot})
    odd = plan_aa(methods, lay, vel, sym)
    assert odd is not None
    return dom, lay, vel, sym, methods, odd


def _run_slabs(dico, size, nsteps, f0_global, slab_axis):
    slabs = [_slab_setup(dico, r, size) for r in range(size)]
    dom0, lay0, vel, sym = slabs[0][0], slabs[0][1], slabs[0][2], slabs[0][3]
    q = len(vel)
    dim = lay0.dim
    vel3 = np.zeros((q, 3), dtype=int)
    vel3[:, 3 - dim:] = vel[:, :dim]
    a = slab_axis                                  # canonical index of the slab axis
    w = lay0.canonical_vmax
    plus = [k for k in range(q) if vel3[k][a] > 0]
    minus = [k for k in range(q) if vel3[k][a] < 0]
    # local arrays: interior of each slab cut out of the global initial state (+ ghosts, any content)
    arrays, offs = [], []
    for dom, lay, *_ in slabs:
        n = lay.canonical_n
        lo = dom.region[0][0]
        A = np.zeros((q,) + tuple(n))
        sl_g = [slice(None)] * 4
        sl_g[1 + a] = slice(lo, lo + n[a])         # global halo array has w ghost layers, same offset
        A[...] = f0_global[tuple(sl_g)]
        arrays.append(A)
        offs.append(lo)

    def planes(A, lo, hi):
        sl = [slice(None)] * 4
        sl[1 + a] = slice(lo, hi)
        return tuple(sl)

    def forward():
        for r, A in enumerate(arrays):
            n = A.shape[1 + a]
            left, right = arrays[(r - 1) % size], arrays[(r + 1) % size]
            nl, nr = left.shape[1 + a], right.shape[1 + a]
            # my planes [w, 2w) -> left neighbour's high ghost (populations moving in -axis)
            for k in minus:
                left[(k,) + planes(left, nl - w[a], nl)[1:]] = A[(k,) + planes(A, w[a], 2 * w[a])[1:]]
            for k in plus:
                right[(k,) + planes(right, 0, w[a])[1:]] = A[(k,) + planes(A, n - 2 * w[a], n - w[a])[1:]]

    def reverse():
        snap = [A.copy() for A in arrays]
        for r, A in enumerate(snap):
            n = A.shape[1 + a]
            left, right = arrays[(r - 1) % size], arrays[(r + 1) % size]
            nl, nr = left.shape[1 + a], right.shape[1 + a]
            # my LOW ghost planes (slots of the populations moving in +axis) -> left neighbour's interior [n-2w, n-w)
            for k in plus:
                left[(k,) + planes(left, nl - 2 * w[a], nl - w[a])[1:]] = A[(k,) + planes(A, 0, w[a])[1:]]
            for k in minus:
                right[(k,) + planes(right, w[a], 2 * w[a])[1:]] = A[(k,) + planes(A, n - w[a], n)[1:]]

    other = tuple(b for b in range(3) if b != a)
    natural = True
    for _ in range(nsteps):
        if natural:
            forward()
            for (dom, lay, _, _, methods, odd), A in zip(slabs, arrays):
                n = lay.canonical_n
                emu._periodic(A, w, other)
                for m in methods:
                    emu._apply(A, m)
                idx = np.meshgrid(*[np.arange(w[b], n[b] - w[b]) for b in range(3)], indexing="ij")
                pulled = [A[(k,) + tuple(slice(w[b] - vel3[k][b], n[b] - w[b] - vel3[k][b]) for b in range(3))].copy()
                          for k in range(q)]
                new = _collide(pulled)
                for k in range(q):
                    tgt = [idx[b] + vel3[k][b] for b in range(3)]
                    A[(sym[k],) + tuple(tgt)] = new[k]
                    out = np.zeros(tgt[0].shape, dtype=bool)
                    wrapped = []
                    for b in range(3):
                        nin = n[b] - 2 * w[b]
                        o = ((tgt[b] < w[b]) | (tgt[b] >= n[b] - w[b])) if (b != a and w[b] > 0) else np.zeros(tgt[b].shape, bool)
                        out |= o
                        wrapped.append(np.where(o, (tgt[b] - w[b]) % max(nin, 1) + w[b], tgt[b]))
                    if out.any():
                        A[(sym[k],) + tuple(t[out] for t in wrapped)] = new[k][out]
            reverse()
            natural = False
        else:
            for (dom, lay, _, _, methods, odd), A in zip(slabs, arrays):
                n = lay.canonical_n
                inner = tuple(slice(w[b], n[b] - w[b]) for b in range(3))
                for m in odd:
                    emu._apply(A, m)
                pulled = [A[(sym[k],) + inner].copy() for k in range(q)]
                new = _collide(pulled)
                for k in range(q):
                    A[(k,) + inner] = new[k]
            natural = True
    # natural view, glued along the slab axis
    parts = []
    for (dom, lay, *_), A in zip(slabs, arrays):
        n = lay.canonical_n
        inner = tuple(slice(w[b], n[b] - w[b]) for b in range(3))
        if natural:
            parts.append(A[(slice(None),) + inner])
        else:
            idx = np.meshgrid(*[np.arange(w[b], n[b] - w[b]) for b in range(3)], indexing="ij")
            S = np.stack([A[(sym[k],) + tuple(idx[b] + vel3[k][b] for b in range(3))] for k in range(q)])
            parts.append(S)
    return np.concatenate(parts, axis=1 + a)


@pytest.mark.parametrize("name", ["cavity2d_bb", "periodic_x_walls_y", "channel2d_inlet_outlet_obstacle",
                                  "cavity3d_bb", "channel3d_d3q27", "bouzidi_walls", "neumann_top"])
@pytest.mark.parametrize("size", [2, 3])
def test_in_place_streaming_on_slabs(name, size):
    dico, _ = emu._directed_cases()[name]
    dom, lay, vel, sym, methods, _ = emu._setup(dico)
    rng = np.random.default_rng(17)
    q = len(vel)
    n, w = lay.canonical_n, lay.canonical_vmax
    slab_axis = 3 - lay.dim
    if n[slab_axis] - 2 * w[slab_axis] < 2 * size * max(1, w[slab_axis]):
        pytest.skip("too few planes for %d slabs" % size)
    f0 = 1.0 / q + 0.05 * rng.uniform(-1, 1, size=(q,) + tuple(n))
    inner = (slice(None),) + tuple(slice(w[i], n[i] - w[i]) for i in range(3))
    for nsteps in (4, 5):
        ref = emu._run(dom, lay, vel, sym, methods, None, nsteps, f0)
        got = _run_slabs(dico, size, nsteps, f0, slab_axis)
        assert got.shape == ref[inner].shape
        assert np.array_equal(got, ref[inner]), (name, size, nsteps)
