"""Stand-alone operators of the C ABI called directly through ctypes on the GPU:
lbm_array_h2d/d2h (padded SoA round trip), lbm_periodic (all populations, both sides) and
lbm_bc_apply (every kind, one- and two-phase) against NumPy."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(nv, shape, vmax, storage="f64"):
    from pylbm_b200.storage import DeviceArray

    return DeviceArray(nv, shape, vmax, storage)


@pytest.mark.parametrize("shape,vmax", [((37,), (2,)), ((19, 23), (1, 1)), ((9, 12, 7), (1, 2, 1))])
@pytest.mark.parametrize("storage", ["f64", "f32"])
def test_padded_round_trip_and_periodic(shape, vmax, storage):
    from pylbm_b200 import runtime as rt

    nv = 5
    rng = np.random.default_rng(0)
    host = rng.uniform(size=(nv,) + shape)
    if storage == "f32":
        host = host.astype(np.float32).astype(np.float64)
    arr = _dev(nv, shape, vmax, storage)
    arr.set(host)
    assert np.array_equal(arr.get(), host)
    assert np.array_equal(arr.get(2, 2), host[2:4])
    # reference semantics of the ghost update, axis by axis (storage.py:333-367 on one rank)
    expect = host.copy()
    for d, w in enumerate(vmax):
        n = shape[d]
        lo = [slice(None)] * (len(shape) + 1)
        src = list(lo)
        lo[d + 1], src[d + 1] = slice(0, w), slice(n - 2 * w, n - w)
        expect[tuple(lo)] = expect[tuple(src)]
        hi, src = [slice(None)] * (len(shape) + 1), [slice(None)] * (len(shape) + 1)
        hi[d + 1], src[d + 1] = slice(n - w, n), slice(w, 2 * w)
        expect[tuple(hi)] = expect[tuple(src)]
    cv = (ctypes.c_int * 3)(*arr.canonical_vmax)
    mask = sum(1 << a for a in range(3) if arr.canonical_vmax[a] > 0)
    rt.check(rt.lib().lbm_periodic(arr.ptr, ctypes.byref(arr.grid), nv, arr.storage_id, cv, mask, None), "lbm_periodic")
    rt.check(rt.lib().lbm_device_sync(), "sync")
    assert np.array_equal(arr.get(), expect)


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("two_phase", [0, 1])
def test_boundary_kernels_bit_exact(kind, two_phase):
    from pylbm_b200 import runtime as rt

    lib = rt.lib()
    nv, shape = 9, (16, 18)
    rng = np.random.default_rng(kind)
    host = rng.uniform(size=(nv,) + shape)
    arr = _dev(nv, shape, (1, 1))
    arr.set(host)
    n = 200
    cells = rng.choice(nv * shape[0] * shape[1], size=3 * n, replace=False)
    idx = np.array(np.unravel_index(cells, (nv,) + shape))      # (3, 3n): disjoint store/load positions
    store, l0, l1 = idx[:, :n], idx[:, n:2 * n], idx[:, 2 * n:]
    rhs, dist = rng.uniform(size=n), rng.uniform(size=n)

    def up(a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        p = ctypes.c_void_p()
        rt.check(lib.lbm_malloc(ctypes.byref(p), a.nbytes), "malloc")
        rt.check(lib.lbm_memcpy_h2d(p, a.ctypes.data, a.nbytes), "h2d")
        return p

    ps, p0, p1 = (up(arr.positions(x), np.int64) for x in (store, l0, l1))
    prhs, pdist, scratch = up(rhs, np.float64), up(dist, np.float64), up(np.zeros(n), np.float64)
    rt.check(lib.lbm_bc_apply(kind, arr.ptr, arr.storage_id, n, ps, p0, p1, prhs, pdist, scratch, two_phase, None),
             "lbm_bc_apply")
    rt.check(lib.lbm_device_sync(), "sync")
    a, b = host[tuple(l0)], host[tuple(l1)]
    expect = host.copy()
    # association order of the reference's generated C (no FMA)
    value = {0: a + rhs, 1: -a + rhs, 2: ((1 - dist) * b + dist * a) + rhs, 3: ((1 - dist) * b - dist * a) + rhs, 4: a}[kind]
    expect[tuple(store)] = value
    assert np.array_equal(arr.get(), expect)
    for p in (ps, p0, p1, prhs, pdist, scratch):
        lib.lbm_free(p)
