"""
ctypes binding of the C ABI in include/lbm_b200.h and include/lbmk.h.

Thin by design: no arithmetic happens here.  Loading fails loudly when the CUDA
runtime library cannot be built/loaded or no GPU is present -- there is no CPU
fallback path in the product.
"""

import ctypes
import json
import os
from ctypes import (
    POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p,
)

from . import build

__all__ = ["LbmkGrid", "LbmkWalls", "LbmSimDesc", "lib", "check", "LbmError", "KernelLibrary", "ensure_gpu"]

STORAGE_F64, STORAGE_F32 = 0, 1
BC_BOUNCE_BACK, BC_ANTI_BOUNCE_BACK, BC_BOUZIDI_BOUNCE_BACK, BC_BOUZIDI_ANTI_BOUNCE_BACK, BC_NEUMANN = range(5)


class LbmError(RuntimeError):
    pass


class LbmkGrid(Structure):
    _fields_ = [
        ("n", c_int * 3),
        ("lo", c_int * 3),
        ("hi", c_int * 3),
        ("tx", c_int),
        ("w", c_int * 3),
        ("wrap", c_int),
        ("pitch", c_int64),
        ("lead", c_int64),
        ("pstride", c_int64),
    ]

    def copy(self):
        out = LbmkGrid()
        ctypes.memmove(ctypes.byref(out), ctypes.byref(self), ctypes.sizeof(LbmkGrid))
        return out


class LbmkWalls(Structure):
    """include/lbmk.h: lbmk_walls"""
    _fields_ = [("lo_plane", c_int), ("hi_plane", c_int), ("neg_lo", c_int), ("neg_hi", c_int),
                ("rhs", c_double * 64)]


LAUNCH_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, POINTER(LbmkGrid), POINTER(c_double), c_void_p)


class LbmSimDesc(Structure):
    _fields_ = [
        ("nv", c_int),
        ("storage", c_int),
        ("grid", LbmkGrid),
        ("vmax", c_int * 3),
        ("periodic_mask", c_int),
        ("f", c_void_p),
        ("fnew", c_void_p),
        ("one_time_step", c_void_p),
        ("one_time_step_peers", c_void_p),
        ("nscalars", c_int),
        ("t_index", c_int),
        ("scalars", c_double * 32),
        ("t", c_double),
        ("dt", c_double),
        ("vel", (ctypes.c_int8 * 3) * 64),
    ]


_lib = None

_SIGNATURES = {
    "lbm_abi_version": (c_int, []),
    "lbm_last_error": (c_char_p, []),
    "lbm_device_count": (c_int, []),
    "lbm_set_device": (c_int, [c_int]),
    "lbm_device_sync": (c_int, []),
    "lbm_mem_info": (c_int, [POINTER(c_uint64), POINTER(c_uint64)]),
    "lbm_malloc": (c_int, [POINTER(c_void_p), c_uint64]),
    "lbm_free": (c_int, [c_void_p]),
    "lbm_memset": (c_int, [c_void_p, c_int, c_uint64]),
    "lbm_host_alloc": (c_int, [POINTER(c_void_p), c_uint64]),
    "lbm_host_free": (c_int, [c_void_p]),
    "lbm_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_uint64]),
    "lbm_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_uint64]),
    "lbm_memcpy_d2d": (c_int, [c_void_p, c_void_p, c_uint64]),
    "lbm_array_h2d": (c_int, [c_void_p, c_void_p, POINTER(LbmkGrid), c_int, c_int, c_int]),
    "lbm_array_d2h": (c_int, [c_void_p, c_void_p, POINTER(LbmkGrid), c_int, c_int, c_int]),
    "lbm_periodic": (c_int, [c_void_p, POINTER(LbmkGrid), c_int, c_int, POINTER(c_int), c_int, c_void_p]),
    "lbm_bc_apply": (
        c_int,
        [c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    ),
    "lbm_sim_create": (c_void_p, [POINTER(LbmSimDesc)]),
    "lbm_sim_destroy": (None, [c_void_p]),
    "lbm_sim_add_bc": (
        c_int,
        [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    ),
    "lbm_sim_set_rhs": (c_int, [c_void_p, c_int, c_void_p]),
    "lbm_sim_bc_groups": (c_int, [c_void_p, c_int, POINTER(c_int)]),
    "lbm_sim_bc_stale_only": (c_int, [c_void_p, c_int, c_int]),
    "lbm_sim_set_walls": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lbm_sim_set_tasks": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
    "lbm_sim_upload_rows": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, c_uint64, c_uint64]),
    "lbm_sim_rhs_update": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_double, c_void_p]),
    "lbm_sim_set_aa": (c_int, [c_void_p, c_void_p]),
    "lbm_sim_set_bc_odd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "lbm_sim_aa_phase": (c_int, [c_void_p]),
    "lbm_sim_set_aa_walls": (c_int, [c_void_p, c_void_p, c_void_p]),
    "lbm_sim_set_scalars": (c_int, [c_void_p, POINTER(c_double), c_int]),
    "lbm_sim_step": (c_int, [c_void_p, c_int]),
    "lbm_sim_boundary_condition": (c_int, [c_void_p]),
    "lbm_sim_sync": (c_int, [c_void_p]),
    "lbm_sim_state": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_double), POINTER(c_int64)]),
    "lbm_sim_set_state": (c_int, [c_void_p, c_void_p, c_void_p, c_double]),
    "lbm_sim_invalidate_ghosts": (c_int, [c_void_p]),
    "lbm_sim_use_graph": (c_int, [c_void_p, c_int]),
    "lbm_sim_timer_start": (c_int, [c_void_p]),
    "lbm_sim_timer_stop": (c_int, [c_void_p, POINTER(c_float)]),
    "lbm_sim_profile": (c_int, [c_void_p, c_int]),
    "lbm_sim_profile_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64)]),
    "lbm_sim_launch_count": (c_int64, [c_void_p]),
    "lbm_sim_stream": (c_void_p, [c_void_p]),
    "lbm_comm_unique_id": (c_int, [c_void_p]),
    "lbm_sim_comm_init": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "lbm_sim_ipc_export": (c_int, [c_void_p, c_void_p]),
    "lbm_sim_ipc_open": (c_int, [c_void_p, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """the loaded runtime library (built on first use)."""
    global _lib
    if _lib is None:
        handle = build.load_library(build.build_runtime())
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.lbm_abi_version() != 1:
            raise LbmError("liblbm_b200.so has ABI version %d, expected 1" % handle.lbm_abi_version())
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc is not None and rc < 0:
        msg = lib().lbm_last_error().decode(errors="replace")
        raise LbmError("%s failed (%d): %s" % (what or "runtime call", rc, msg))
    return rc


def ensure_gpu():
    n = lib().lbm_device_count()
    if n <= 0:
        raise LbmError(
            "no CUDA device available (%s); generator='cuda' has no CPU fallback"
            % lib().lbm_last_error().decode(errors="replace")
        )
    return n


class _PinnedBlock:
    """page-locked host memory exposed through the array interface.  Page-locking is slow (~0.5 s per
    GB) while a DMA into page-locked memory is ~5x faster than a copy into pageable memory, so blocks
    go back to a small pool when their last view dies and repeated reads of fields of the same size
    (`sol.m[...]` every few steps) reuse them."""

    _pool = {}          # nbytes -> [ptr, ...]
    _pooled_bytes = 0
    try:                # at most 16 GB, and never more than a quarter of the host memory
        POOL_LIMIT = min(16 << 30, os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") // 4)
    except (ValueError, OSError, AttributeError):
        POOL_LIMIT = 4 << 30

    def __init__(self, count):
        self.nbytes = int(count) * 8
        free = _PinnedBlock._pool.get(self.nbytes)
        if free:
            self.ptr = free.pop()
            _PinnedBlock._pooled_bytes -= self.nbytes
        else:
            ptr = c_void_p()
            check(lib().lbm_host_alloc(ctypes.byref(ptr), self.nbytes), "lbm_host_alloc")
            self.ptr = ptr.value
        self.__array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if not ptr:
            return
        try:
            if _PinnedBlock._pooled_bytes + self.nbytes <= _PinnedBlock.POOL_LIMIT:
                _PinnedBlock._pool.setdefault(self.nbytes, []).append(ptr)
                _PinnedBlock._pooled_bytes += self.nbytes
            else:
                lib().lbm_host_free(ptr)
        except Exception:
            pass


PINNED_MIN_BYTES = 8 << 20


def host_empty(shape):
    """float64 host array for a device -> host copy: page-locked (pooled) when it is large enough
    for the copy speed to matter, plain NumPy otherwise."""
    import numpy as np

    count = 1
    for n in shape:
        count *= int(n)
    if count * 8 < PINNED_MIN_BYTES:
        return np.empty(shape, dtype=np.float64)
    try:
        return np.asarray(_PinnedBlock(count)).reshape(shape)
    except LbmError:
        return np.empty(shape, dtype=np.float64)


class KernelLibrary:
    """a generated per-scheme library (include/lbmk.h)."""

    def __init__(self, path):
        self.path = path
        self.handle = build.load_library(path)
        self.handle.lbmk_abi_version.restype = c_int
        self.handle.lbmk_describe.restype = c_char_p
        if self.handle.lbmk_abi_version() != 1:
            raise LbmError("%s: unexpected kernel ABI version" % path)
        self.info = json.loads(self.handle.lbmk_describe().decode())
        self.routines = {}
        for name in self.info["routines"]:
            fn = getattr(self.handle, "lbmk_" + name)
            fn.restype = c_int
            fn.argtypes = [c_void_p, c_void_p, POINTER(LbmkGrid), POINTER(c_double), c_void_p]
            self.routines[name] = fn

    def scalars(self, name):
        return self.info["routines"][name]["scalars"]

    def address(self, name):
        if name not in self.routines:
            fn = getattr(self.handle, "lbmk_" + name)
            return ctypes.cast(fn, c_void_p).value
        return ctypes.cast(self.routines[name], c_void_p).value

    def launch(self, name, fin, fout, grid, scalars=(), stream=None):
        arr = (c_double * max(1, len(scalars)))(*scalars)
        rc = self.routines[name](fin, fout, ctypes.byref(grid), arr, stream)
        if rc != 0:
            raise LbmError("kernel %s failed to launch (cudaError %d)" % (name, -rc))
