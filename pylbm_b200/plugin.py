"""
Registration of `generator='cuda'` inside an importable reference pylbm.

pylbm selects a backend by the string `dico['generator']` in four hard-coded tables
(SURVEY.md 8b).  `register()` extends them at run time -- the reference tree is never edited:

  1. validator          pylbm/validator.py:327-329      'cuda' admitted
  2. container table    pylbm/simulation.py:165-171     "CUDA" -> CudaRefContainer (HBM-resident arrays)
  3. code generator     pylbm/generator/codegen.py:1484-1494   "CUDA" -> CudaCodeGen
  4. code wrapper       pylbm/generator/autowrap.py:155-163    "CUDA" -> CudaCodeWrapper

After that the UNCHANGED `pylbm.Simulation(dico)` drives the CUDA kernels: its symbolic algorithm
(`pylbm/algorithm/base.py:602-628` adds one `Routine` per kernel, a `For` over `Eq` statements on
`Indexed` arrays) is lowered by `routine_to_ir()` into this package's per-cell IR and compiled by
cudagen/nvcc; the boundary routines (`bounce_back`, `Bouzidi_bounce_back`, ...) are mapped by name to
the runtime's boundary kernels.  The module-like object returned to the driver exposes one callable
per routine with an `arg_dict` (the kwargs protocol of `pylbm/symbolic.py:288-299`).

This is the compatibility path: every piece of the reference's per-step Python still runs (ghost
update, set_rhs, one call per boundary method, kernel, swap).  The fast path -- one runtime call per
step -- is `pylbm_b200.Simulation`.

Status: the IR lowering is tested on the CPU against the CPU checker whenever the reference is importable
(tests/test_plugin_ir.py); the device classes need pylbm AND a GPU in the same process, which the
GPU box of this project does not offer (no pylbm there), so they are exercised only to the point of
the first device allocation.
"""

import ctypes

import numpy as np
import sympy as sp

from . import runtime as rt
from .algorithm import KernelIR

__all__ = ["register", "routine_to_ir", "BC_ROUTINES"]

BC_ROUTINES = {
    "bounce_back": rt.BC_BOUNCE_BACK,
    "anti_bounce_back": rt.BC_ANTI_BOUNCE_BACK,
    "Bouzidi_bounce_back": rt.BC_BOUZIDI_BOUNCE_BACK,
    "Bouzidi_anti_bounce_back": rt.BC_BOUZIDI_ANTI_BOUNCE_BACK,
    "neumann": rt.BC_NEUMANN,
    "neumannx": rt.BC_NEUMANN,
    "neumanny": rt.BC_NEUMANN,
    "neumannz": rt.BC_NEUMANN,
}
_ARRAYS = ("f", "fnew", "m")
_LOOP = {"ix_": 0, "iy_": 1, "iz_": 2}


# ---------------------------------------------------------------------------
# reference Routine -> KernelIR
# ---------------------------------------------------------------------------
def _split_index(indexed):
    """Indexed f[...] -> (population k, offset per space axis); the population slot is the integer
    index, the others are `loop symbol + integer` whatever the storage order (symbolic.py:113-208)."""
    k, offset = None, {}
    for idx in indexed.indices:
        idx = sp.sympify(idx)
        if idx.is_Integer:
            if k is not None:
                raise NotImplementedError("two integer indices in %s" % indexed)
            k = int(idx)
            continue
        syms = [s for s in idx.free_symbols if s.name in _LOOP]
        if len(syms) != 1:
            raise NotImplementedError("unsupported index %s in %s" % (idx, indexed))
        shift = sp.simplify(idx - syms[0])
        if not shift.is_Integer:
            raise NotImplementedError("non-constant shift in %s" % indexed)
        offset[_LOOP[syms[0].name]] = int(shift)
    if k is None:
        raise NotImplementedError("no population index in %s" % indexed)
    return k, tuple(offset[a] for a in sorted(offset))


def routine_to_ir(routine):
    """
    Lower one per-cell Routine of the reference (`For(space indices, [Eq, ...])`) to a KernelIR.
    Matrix equations are element-wise sequential in-place assignments, as the reference prints them
    (generator/printing/cython.py:316-350).
    """
    from sympy.matrices.expressions.matexpr import MatrixElement

    (loop,) = routine.statements
    body = list(loop.body.args) if hasattr(loop.body, "args") else list(loop.body)
    inner = any(sp.sympify(i.lower) != 0 for i in loop.target)
    dim = len(loop.target)

    reads = {}      # (array, k) -> (symbol, offset)
    local = {}      # current symbol of an array cell written by this kernel: (array, k) -> symbol
    written = {}    # (array, k) -> True
    statements = []

    def cell_symbol(array, k):
        return sp.Symbol("%s%d" % (array, k), real=True)

    def convert(expr):
        expr = sp.sympify(expr)
        repl = {}
        for node in expr.atoms(sp.Indexed):
            name = str(node.base)
            if name not in _ARRAYS:
                raise NotImplementedError("array %s in a per-cell kernel" % name)
            k, off = _split_index(node)
            zero = all(o == 0 for o in off)
            if zero and (name, k) in written:
                repl[node] = cell_symbol(name, k)        # reads what this kernel stored (sequential)
                continue
            key = (name, k)
            if key in reads and reads[key][1] != off:
                raise NotImplementedError("population %d of %s read at two offsets" % (k, name))
            reads[key] = (cell_symbol(name, k), off)
            repl[node] = reads[key][0]
        for node in expr.atoms(MatrixElement):
            repl[node] = sp.Symbol("%s_l%d" % (node.parent, int(node.i)), real=True)
        return expr.xreplace(repl)

    def assign(lhs, rhs):
        rhs = convert(rhs)
        if isinstance(lhs, sp.Indexed):
            name = str(lhs.base)
            k, off = _split_index(lhs)
            if any(o != 0 for o in off):
                raise NotImplementedError("store with an offset: %s" % lhs)
            written[(name, k)] = True
            statements.append((cell_symbol(name, k), rhs))
        elif isinstance(lhs, MatrixElement):
            statements.append((sp.Symbol("%s_l%d" % (lhs.parent, int(lhs.i)), real=True), rhs))
        elif isinstance(lhs, sp.Symbol):
            statements.append((lhs, rhs))
        else:
            raise NotImplementedError("left-hand side %r" % lhs)

    for eq in body:
        lhs, rhs = eq.lhs, eq.rhs
        if hasattr(lhs, "shape") and not isinstance(lhs, (sp.Indexed, MatrixElement)):
            rows, cols = lhs.shape
            for i in range(rows):
                for j in range(cols):
                    if lhs[i, j] == rhs[i, j]:
                        continue                       # `m[i] = m[i]` is not printed by the reference
                    assign(lhs[i, j], rhs[i, j])
        else:
            assign(lhs, rhs)

    in_arrays = sorted({a for a, _ in reads})
    out_arrays = sorted({a for a, _ in written})
    if len(in_arrays) != 1 or len(out_arrays) != 1:
        raise NotImplementedError("kernel %s reads %s and writes %s" % (routine.name, in_arrays, out_arrays))
    in_array, out_array = in_arrays[0], out_arrays[0]
    nq = 1 + max(max(k for _, k in reads), max(k for _, k in written))
    in_syms, in_offsets = [], []
    for k in range(nq):
        sym, off = reads.get((in_array, k), (sp.Symbol("%s%d" % (in_array, k), real=True), (0,) * dim))
        in_syms.append(sym)
        in_offsets.append(off)
    outputs = []
    for k in range(nq):
        if (out_array, k) in written:
            outputs.append(cell_symbol(out_array, k))
        elif in_array == out_array:
            outputs.append(in_syms[k])                 # untouched entry of an in-place kernel
        else:
            raise NotImplementedError("population %d of %s is never stored" % (k, out_array))
    # a kernel that reads and writes the same array in place: the loaded and the stored symbol of a
    # cell have the same name, which is exactly the sequential semantics (SSA renames later)
    known = set(in_syms) | {lhs for lhs, _ in statements}
    free = set()
    for _, rhs in statements:
        free |= rhs.free_symbols
    scalars = sorted(str(s) for s in free - known)
    return KernelIR(routine.name, in_array, in_syms, in_offsets, out_array, statements, outputs, inner, scalars)


# ---------------------------------------------------------------------------
# device arrays seen through the reference's Array interface
# ---------------------------------------------------------------------------
class _Handle:
    """what the driver passes around as `Array.array` (storage.py:107-118)."""

    def __init__(self, dev):
        self.dev = dev

    def __getitem__(self, key):
        return self

    def __setitem__(self, key, other):      # Fnew.array[:] = F.array[:]  (simulation.py:320)
        if isinstance(other, _Handle):
            self.dev.copy_from(other.dev)
        else:
            self.dev.set(np.asarray(other))

    def copy(self):                          # fcopy = F.array.copy()  (boundary.py:549): the Bouzidi
        return self                          # kernel gathers before it scatters, no snapshot needed

    shape = property(lambda self: self.dev.shape)
    size = property(lambda self: self.dev.size)


class CudaRefArray:
    """pylbm.storage.Array look-alike over a padded SoA DeviceArray."""

    gpu_support = False      # keeps the reference away from its pyopencl branches (storage.py:109-157)

    def __init__(self, nv, shape_halo, vmax, consm=None):
        from .storage import DeviceArray

        self.dev = DeviceArray(nv, shape_halo, vmax, "f64", consm)
        self.array = _Handle(self.dev)
        self.vmax = list(vmax)
        self.consm = self.dev.consm
        self.sorder = self.index = list(range(len(shape_halo) + 1))
        self.dim = len(shape_halo)

    nspace = property(lambda self: self.dev.nspace)
    nv = property(lambda self: self.dev.nv)
    shape = property(lambda self: self.dev.shape)
    size = property(lambda self: self.dev.size)
    swaparray = property(lambda self: self.dev.get())

    def set_conserved_moments(self, consm):
        self.dev.set_conserved_moments(consm)

    def __getitem__(self, key):
        return self.dev[key]

    def __setitem__(self, key, values):
        self.dev[key] = values

    def _in(self, key):
        return self.dev._in(key)

    def generate(self, generator):
        pass

    def update(self):
        """ghost update of one rank (storage.py:306-367)."""
        vmax = (ctypes.c_int * 3)(*self.dev.canonical_vmax)
        mask = sum(1 << a for a in range(3) if self.dev.canonical_vmax[a] > 0)
        rt.check(
            rt.lib().lbm_periodic(self.dev.ptr, ctypes.byref(self.dev.grid), self.dev.nv, self.dev.storage_id,
                                  vmax, mask, None),
            "lbm_periodic",
        )


class CudaRefContainer:
    """the 'CUDA' entry of simulation.py:165-171."""

    gpu_support = False

    def __init__(self, domain, scheme, sorder=None):
        self.dim = domain.dim
        self.mpi_topo = domain.mpi_topo
        self.nv = int(scheme.stencil.nv_ptr[-1])
        self.nspace = domain.global_size
        self.vmax = list(domain.stencil.vmax)
        self.sorder = list(range(self.dim + 1))
        shape = domain.shape_halo
        self.m = CudaRefArray(self.nv, shape, self.vmax, scheme.consm)
        self.F = CudaRefArray(self.nv, shape, self.vmax, scheme.consm)
        self.Fnew = CudaRefArray(self.nv, shape, self.vmax, scheme.consm)

    def move2gpu(self, array):
        return array


# ---------------------------------------------------------------------------
# module-like object returned to the driver
# ---------------------------------------------------------------------------
class CudaModule:
    def __init__(self, routines, dim_hint=None):
        from . import build
        from .cudagen import generate_source

        kernels, self._bc = [], {}
        for r in routines:
            if r.name in BC_ROUTINES:
                self._bc[r.name] = BC_ROUTINES[r.name]
            else:
                kernels.append(routine_to_ir(r))
        self._device_lists = {}
        self._scratch = None
        if kernels:
            dim = len(kernels[0].in_offsets[0])
            nv = len(kernels[0].in_syms)
            source, info = generate_source(kernels, dim, nv)
            self.library = rt.KernelLibrary(build.build_kernels(source, info["hash"]))
            self.source = source
            for ir in kernels:
                setattr(self, ir.name, self._kernel(ir))
        for name, kind in self._bc.items():
            setattr(self, name, self._boundary(name, kind))

    # ---- per-cell kernels ----
    def _kernel(self, ir):
        lib, scalars = self.library, list(ir.scalars)
        in_name, out_name = ir.in_array, ir.out_array

        def call(queue=None, **kw):
            src, dst = kw[in_name], kw[out_name]
            values = [float(kw[s]) for s in scalars]
            if isinstance(src, _Handle):
                grid = src.dev.inner_grid() if ir.inner else src.dev.grid
                lib.launch(ir.name, src.dev.ptr, dst.dev.ptr, grid, values)
                rt.check(rt.lib().lbm_device_sync(), "sync")
                return
            # small host arrays [nv, n, 1(, 1)] (wall equilibria: boundary.py:275-293)
            from .storage import DeviceArray

            ncell = int(np.prod(src.shape[1:]))
            dsrc = DeviceArray(src.shape[0], (ncell,), [0], "f64")
            dsrc.set(np.ascontiguousarray(src).reshape(src.shape[0], ncell))
            ddst = dsrc if dst is src else DeviceArray(dst.shape[0], (ncell,), [0], "f64")
            lib.launch(ir.name, dsrc.ptr, ddst.ptr, dsrc.grid, values)
            rt.check(rt.lib().lbm_device_sync(), "sync")
            dst[...] = ddst.get().reshape(dst.shape)

        call.arg_dict = {k: None for k in dict.fromkeys([in_name, out_name] + scalars)}
        call.__name__ = ir.name
        return call

    # ---- boundary kernels ----
    def _positions(self, handle, index):
        key = (index.ctypes.data, index.shape)
        if key not in self._device_lists:
            pos = np.ascontiguousarray(handle.dev.positions(index.T))
            ptr = ctypes.c_void_p()
            rt.check(rt.lib().lbm_malloc(ctypes.byref(ptr), max(8, pos.nbytes)), "lbm_malloc")
            rt.check(rt.lib().lbm_memcpy_h2d(ptr, pos.ctypes.data, pos.nbytes), "h2d")
            self._device_lists[key] = ptr.value
        return self._device_lists[key]

    def _upload(self, array):
        array = np.ascontiguousarray(array, dtype=np.float64)
        ptr = ctypes.c_void_p()
        rt.check(rt.lib().lbm_malloc(ctypes.byref(ptr), max(8, array.nbytes)), "lbm_malloc")
        rt.check(rt.lib().lbm_memcpy_h2d(ptr, array.ctypes.data, array.nbytes), "h2d")
        return ptr.value

    def _boundary(self, name, kind):
        two_loads = kind in (rt.BC_BOUZIDI_BOUNCE_BACK, rt.BC_BOUZIDI_ANTI_BOUNCE_BACK)
        names = ["f", "istore", "iload0", "ncond"] + (["iload1", "dist"] if two_loads else [])
        if kind != rt.BC_NEUMANN:
            names.append("rhs")
        if kind == rt.BC_BOUZIDI_BOUNCE_BACK:
            names.append("fcopy")

        def call(queue=None, **kw):
            f, ncond = kw["f"], int(kw["ncond"])
            if ncond == 0:
                return
            store = self._positions(f, kw["istore"])
            l0 = self._positions(f, kw["iload0"])
            l1 = self._positions(f, kw["iload1"]) if two_loads else None
            rhs = self._upload(kw["rhs"]) if kind != rt.BC_NEUMANN else None
            dist = self._upload(kw["dist"]) if two_loads else None
            scratch = self._upload(np.zeros(ncond))
            # the reference loop is sequential: gather-then-scatter keeps its result unless an entry
            # reads what an EARLIER entry of the same call stored (see boundary.schedule); the
            # stand-alone Simulation handles that case with levels
            rc = rt.lib().lbm_bc_apply(kind, f.dev.ptr, f.dev.storage_id, ncond, store, l0, l1, rhs, dist,
                                       scratch, 1, None)
            rt.check(rc, "lbm_bc_apply(%s)" % name)
            rt.check(rt.lib().lbm_device_sync(), "sync")
            for ptr in (rhs, dist, scratch):
                if ptr:
                    rt.lib().lbm_free(ptr)

        call.arg_dict = {k: None for k in names}
        call.__name__ = name
        return call


# ---------------------------------------------------------------------------
# registration
# ---------------------------------------------------------------------------
_registered = False


def register():
    """make `generator='cuda'` known to the importable pylbm (idempotent)."""
    global _registered
    if _registered:
        return
    import importlib

    import pylbm

    # (pylbm.generator re-exports functions named like its sub-modules: go through importlib)
    ref_simulation = importlib.import_module("pylbm.simulation")
    ref_autowrap = importlib.import_module("pylbm.generator.autowrap")
    ref_codegen = importlib.import_module("pylbm.generator.codegen")

    # 3. code generator: the reference's CodeGen.routine() is backend independent
    #    (codegen.py:763-897); only `has_output` and the printer matter to it
    class CudaCodeGen(ref_codegen.CythonCodeGen):
        language = "cuda"

    original_generator = ref_codegen.get_code_generator

    def get_code_generator(language, project=None, standard=None, printer=None):
        if str(language).upper() == "CUDA":
            return CudaCodeGen(project=project)
        return original_generator(language, project, standard, printer)

    ref_codegen.get_code_generator = get_code_generator
    ref_autowrap.get_code_generator = get_code_generator

    # 4. code wrapper (autowrap.py:52-76 contract: wrap_code(routines) -> module-like object)
    class CudaCodeWrapper:
        def __init__(self, generator, filepath=None, flags=(), generate=True, verbose=False):
            self.verbose = verbose

        def wrap_code(self, routines):
            module = CudaModule(list(routines))
            if self.verbose:
                print(module.source)
            return module

    original_wrapper = ref_autowrap.get_code_wrapper

    def get_code_wrapper(backend):
        if str(backend).upper() == "CUDA":
            return CudaCodeWrapper
        return original_wrapper(backend)

    ref_autowrap.get_code_wrapper = get_code_wrapper

    # 2. container table (simulation.py:165-171)
    original_container = ref_simulation.Simulation._get_container

    def _get_container(self, sorder):
        if self.generator.backend == "CUDA":
            rt.ensure_gpu()
            return CudaRefContainer(self.domain, self.scheme, sorder)
        return original_container(self, sorder)

    ref_simulation.Simulation._get_container = _get_container

    # 1. validator (validator.py:327-329): accept the new name when cerberus is the real one
    try:
        ref_validator = importlib.import_module("pylbm.validator")

        original_validate = ref_validator.validate

        def validate(dico, name):
            if str(dico.get("generator", "")).lower() == "cuda":
                patched = dict(dico)
                patched["generator"] = "cython"
                return original_validate(patched, name)
            return original_validate(dico, name)

        ref_validator.validate = validate
        ref_simulation.validate = validate
    except Exception:       # pragma: no cover
        pass
    _registered = True
    return pylbm
