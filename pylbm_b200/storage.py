"""
Device-resident storage of the distribution functions and moments.

Replaces pylbm.storage.Array/SOA/AOS (reference: pylbm/storage.py:24-207,
423-524) for generator='cuda'.  The reference allocates a dense NumPy array
permuted by `sorder` and (for its OpenCL backend) mirrors the whole array to the
device on every access.  Here the array lives in HBM only, as

    [population k][x][y][z]   (structure of arrays, z fastest)

with padded rows: the row pitch is a multiple of 128 bytes and the row origin is
shifted so that the FIRST INTERIOR cell of every row is 128-byte aligned (ghost
cells sit at the end of the previous 128-byte line).  Every warp-wide load/store
of a population in the fused kernel is then one or two full 128-byte lines.
The host sees dense `[nv, x, y, z]` NumPy blocks through explicit, padding-aware
copies (`lbm_array_h2d/d2h`), population by population; `sorder` is accepted for
API compatibility and only affects nothing: the layout is fixed.

HostArray is the small NumPy-backed twin used for boundary-value evaluation
(reference: pylbm/boundary.py:275-290 builds tiny Arrays the user's `value`
callback writes into with `m[rho] = ...`).
"""

import ctypes

import numpy as np
import sympy as sp

from . import runtime as rt

__all__ = ["DeviceArray", "DeviceView", "HostArray"]


def _roundup(n, m):
    return (n + m - 1) // m * m


def block_tx(n_fast):
    """threads of a 128-thread block laid along the fastest axis (power of two): whole rows of short
    lattices share a block (lbmk_grid.tx)."""
    tx = 128
    while tx > 1 and tx // 2 >= n_fast:
        tx //= 2
    return tx


class _ConsmMixin:
    def set_conserved_moments(self, consm):
        for k, v in consm.items():
            self.consm[k] = v

    def _key(self, key):
        if isinstance(key, (sp.Symbol, sp.IndexedBase)):
            return self.consm[key]
        return key


class HostArray(_ConsmMixin):
    """dense host array [nv, *nspace] with conserved-moment keys."""

    def __init__(self, nv, nspace, vmax=None, consm=None):
        self.nspace = tuple(int(n) for n in nspace)
        self.vmax = list(vmax) if vmax is not None else [0] * len(self.nspace)
        self.array = np.zeros((nv,) + self.nspace)
        self.swaparray = self.array
        self.consm = dict(consm or {})

    nv = property(lambda self: self.array.shape[0])
    shape = property(lambda self: self.array.shape)

    def __getitem__(self, key):
        return self.array[self._key(key)]

    def __setitem__(self, key, values):
        self.array[self._key(key)] = values


class Layout:
    """element layout of a padded SoA array (no device memory: usable on the build box)."""

    def __init__(self, nv, nspace, vmax, itemsize=8, align=None):
        self.nv = int(nv)
        self.nspace = tuple(int(n) for n in nspace)
        self.dim = len(self.nspace)
        self.vmax = [int(v) for v in vmax]
        n = (1,) * (3 - self.dim) + self.nspace
        w = (0,) * (3 - self.dim) + tuple(self.vmax)
        align = int(align) if align else 128 // itemsize
        self.align = align
        self.pitch = _roundup(n[2], align)
        self.lead = (align - w[2]) % align
        self.pstride = _roundup(self.lead + n[0] * n[1] * self.pitch, align)
        self.canonical_n, self.canonical_vmax = n, w
        self.tx = block_tx(n[2])

    def positions(self, index):
        """element positions of entries [k, ix(, iy(, iz))] given as an integer array (dim+1, n)."""
        index = np.asarray(index, dtype=np.int64)
        k = index[0]
        space = [np.zeros_like(k)] * (3 - self.dim) + [index[1 + d] for d in range(self.dim)]
        n = self.canonical_n
        return k * self.pstride + self.lead + (space[0] * n[1] + space[1]) * self.pitch + space[2]


class DeviceView:
    """what `Array.array` is for a device array (reference: storage.py:107-118 hands a NumPy or
    pyopencl array to the generated functions): an opaque handle the CUDA module recognises."""

    def __init__(self, dev):
        self.dev = dev

    def __getitem__(self, key):
        return self

    def __setitem__(self, key, other):      # Fnew.array[:] = F.array[:]  (simulation.py:320)
        if isinstance(other, DeviceView):
            self.dev.copy_from(other.dev)
        else:
            self.dev.set(np.asarray(other))

    def copy(self):                          # fcopy = F.array.copy()  (boundary.py:549): the Bouzidi
        return self                          # kernel gathers before it scatters, no snapshot needed

    shape = property(lambda self: self.dev.shape)
    size = property(lambda self: self.dev.size)


class DeviceArray(_ConsmMixin):
    """
    Padded SoA array in HBM.  `nspace` includes the ghost layers (`vmax` per side).
    `storage` is 'f64' (default, the reference's only type: storage.py:67) or 'f32'
    (optional reduced-precision storage, arithmetic stays fp64).
    `align` = row alignment in ELEMENTS (default: 128 bytes of this array's own type).  A kernel
    addresses its input and output arrays with ONE lbmk_grid, so arrays of different element types
    that meet in a kernel (fp32 populations and fp64 moments) must be built with the same `align`.
    """

    def __init__(self, nv, nspace, vmax, storage="f64", consm=None, align=None):
        rt.ensure_gpu()
        self.nv = int(nv)
        self.nspace = tuple(int(n) for n in nspace)
        self.dim = len(self.nspace)
        self.vmax = [int(v) for v in vmax]
        self.storage = storage
        self.storage_id = rt.STORAGE_F64 if storage == "f64" else rt.STORAGE_F32
        self.itemsize = 8 if storage == "f64" else 4
        self.consm = dict(consm or {})
        self.sorder = list(range(self.dim + 1))

        n = (1,) * (3 - self.dim) + self.nspace
        w = (0,) * (3 - self.dim) + tuple(self.vmax)
        align = int(align) if align else 128 // self.itemsize
        self.align = align
        pitch = _roundup(n[2], align)
        lead = (align - w[2]) % align
        rows = n[0] * n[1]
        pstride = _roundup(lead + rows * pitch, align)
        self.canonical_n = n
        self.canonical_vmax = w
        self.pitch, self.lead, self.pstride = pitch, lead, pstride
        self.nbytes = self.nv * pstride * self.itemsize

        self.grid = rt.LbmkGrid()
        for a in range(3):
            self.grid.n[a] = n[a]
            self.grid.lo[a] = 0
            self.grid.hi[a] = n[a]
            self.grid.w[a] = w[a]
        self.grid.wrap = 0
        self.tx = self.grid.tx = block_tx(n[2])
        self.grid.pitch, self.grid.lead, self.grid.pstride = pitch, lead, pstride

        ptr = ctypes.c_void_p()
        rt.check(rt.lib().lbm_malloc(ctypes.byref(ptr), self.nbytes), "lbm_malloc(%d bytes)" % self.nbytes)
        self.ptr = ptr.value
        rt.check(rt.lib().lbm_memset(self.ptr, 0, self.nbytes), "lbm_memset")

    def __del__(self):
        ptr = getattr(self, "ptr", None)
        if ptr:
            try:
                rt.lib().lbm_free(ptr)
            except Exception:
                pass
            self.ptr = None

    # ---- the rest of the reference's Array surface (storage.py:107-118, 306-367) ------------
    array = property(lambda self: DeviceView(self))
    swaparray = property(lambda self: self.get())
    gpu_support = True

    def generate(self, generator):
        """the ghost-update kernels are part of the static runtime (k_periodic): nothing to generate
        (reference: storage.py:370-420 generates update_x/y/z for its OpenCL backend)."""

    def update(self):
        """periodic ghost update of one rank, dimension by dimension (reference: storage.py:306-367)."""
        vmax = (ctypes.c_int * 3)(*self.canonical_vmax)
        mask = sum(1 << a for a in range(3) if self.canonical_vmax[a] > 0)
        rt.check(rt.lib().lbm_periodic(self.ptr, ctypes.byref(self.grid), self.nv, self.storage_id, vmax, mask, None),
                 "lbm_periodic")

    # ---- geometry helpers ------------------------------------------------
    def inner_grid(self):
        """grid whose lo/hi is the interior [vmax, n - vmax) (reference: base.py:170-189)."""
        g = self.grid.copy()
        for a in range(3):
            g.lo[a] = self.canonical_vmax[a]
            g.hi[a] = self.canonical_n[a] - self.canonical_vmax[a]
        return g

    def positions(self, index):
        """
        element positions of entries [k, ix(, iy(, iz))] given as an integer array of
        shape (dim+1, n) (the reference's istore/iload layout before fix_iload).
        """
        index = np.asarray(index, dtype=np.int64)
        k = index[0]
        space = [np.zeros_like(k)] * (3 - self.dim) + [index[1 + d] for d in range(self.dim)]
        n = self.canonical_n
        return k * self.pstride + self.lead + (space[0] * n[1] + space[1]) * self.pitch + space[2]

    @property
    def shape(self):
        return (self.nv,) + self.nspace

    @property
    def size(self):
        return int(np.prod(self.shape))

    # ---- host <-> device ---------------------------------------------------
    def get(self, k0=0, nk=None):
        """dense host copy of populations k0..k0+nk-1: array [nk, *nspace]."""
        nk = self.nv - k0 if nk is None else nk
        host = rt.host_empty((nk,) + self.nspace)
        rt.check(
            rt.lib().lbm_array_d2h(host.ctypes.data, self.ptr, ctypes.byref(self.grid), self.storage_id, k0, nk),
            "lbm_array_d2h",
        )
        return host

    def set(self, values, k0=0):
        """upload a dense host block [nk, *nspace] (or [*nspace]) to populations k0.."""
        values = np.asarray(values, dtype=np.float64)
        if values.ndim == self.dim:
            values = values[np.newaxis]
        if values.shape[1:] != self.nspace:
            raise ValueError("shape mismatch: %s vs %s" % (values.shape[1:], self.nspace))
        values = np.ascontiguousarray(values)
        rt.check(
            rt.lib().lbm_array_h2d(
                self.ptr, values.ctypes.data, ctypes.byref(self.grid), self.storage_id, k0, values.shape[0]
            ),
            "lbm_array_h2d",
        )

    def copy_from(self, other):
        if other.nbytes != self.nbytes:
            raise ValueError("device arrays differ in size")
        rt.check(rt.lib().lbm_memcpy_d2d(self.ptr, other.ptr, self.nbytes), "lbm_memcpy_d2d")

    # ---- reference-style access (host copies) ------------------------------
    def __getitem__(self, key):
        key = self._key(key)
        if isinstance(key, (int, np.integer)):
            return self.get(int(key), 1)[0]
        return self.get()[key]

    def __setitem__(self, key, values):
        key = self._key(key)
        if isinstance(key, (int, np.integer)):
            host = np.empty(self.nspace)
            host[...] = values
            self.set(host, int(key))
        elif (key is Ellipsis or (isinstance(key, slice) and key == slice(None))) \
                and getattr(values, "shape", None) == self.shape:
            self.set(values)                     # whole array: no read-modify-write
        else:
            host = self.get()
            host[key] = values
            self.set(host)

    def _in(self, key):
        inner = tuple(slice(v, -v) if v > 0 else slice(None) for v in self.vmax)
        return self[key][inner]
