"""
`generator='cuda'` inside an importable reference pylbm: the drop-in boundary of SURVEY.md 8(b).

pylbm selects a backend by the string `dico['generator']` in hard-coded tables.  `register()` extends
them at run time -- the reference tree is never edited:

  1. validator          pylbm/validator.py:327-329              'cuda' admitted
  2. container table    pylbm/simulation.py:165-171             "CUDA" -> CudaContainer (padded SoA in HBM)
  3. code generator     pylbm/generator/codegen.py:1484-1494    "CUDA" -> CudaCodeGen
  4. code wrapper       pylbm/generator/autowrap.py:155-163     "CUDA" -> CudaCodeWrapper / Generator.compile
  5. `pylbm.Simulation(dico)` with generator='cuda' instantiates `CudaSimulation`, a SUBCLASS of the
     reference class (one `__new__` hook).  Its constructor IS the reference's constructor
     (simulation.py:89-153, unchanged, called through super()): validate -> Domain -> Scheme -> Generator
     -> container -> algorithm.generate() -> Boundary -> compile -> initialize.  Two names the constructor
     resolves in its module are redirected for this backend: `Domain` builds the sparse records of
     pylbm_b200.domain from the reference's own Geometry / Stencil / elements (the dense
     [unvtot, nx, ny, nz] arrays of domain.py:285-293 are ~105 GB for D3Q19 at 512^3), and `Boundary`
     builds the same ordered index lists from them (bit-identical, tests/test_gpu_plugin.py).

What the subclass overrides is the hot path (CudaEngine): `one_time_step()` is ONE enqueue-only runtime
call (`lbm_sim_step`: ghost update, boundary kernels over device-resident lists uploaded once, fused
pull kernel, swap), `boundary_condition()` one `lbm_sim_boundary_condition`; nothing is allocated,
copied or synchronised per step.  `sol.m[...]`, `sol.F[...]`, `f2m/m2f/equilibrium/relaxation/
transport/source_term`, `sol.t/nt/dt`, `extra_parameters` keep the reference's meaning.

Kernels.  Two lowerings produce the per-scheme CUDA library, both compiled by cudagen/nvcc:
  * 'scheme' (default for the stock PullAlgorithm): the scheme-level lowering of algorithm.py (numeric
    transforms + sparse relative-velocity shifts; 2x fewer fp64 operations than the reference's dense
    polynomial matrices when `relative_velocity` is used);
  * 'ir' (`dico['cuda_option'] = {'lowering': 'ir'}`, and always for a user-defined algorithm class): the
    reference's own symbolic Routines (algorithm/base.py:602-628: a `For` over `Eq` statements on
    `Indexed` arrays) are lowered statement by statement by `routine_to_ir()`.
Either way the module-like object handed back to the reference (`sol.generator.module`) exposes one
callable per routine with an `arg_dict` (the kwargs protocol of pylbm/symbolic.py:288-299), so
`sol.algo.call_function(name, sol)` keeps working.

Multi-GPU: one process per GPU.  When `torch.distributed` is initialised with more than one rank (the
role MPI.COMM_WORLD plays in the reference, mpi_topology.py:74-105) the lattice is cut into x-slabs,
one per rank, with the fused NVLink halo; `configure()` overrides the discovery.
"""

import ctypes

import numpy as np
import sympy as sp

from . import runtime as rt
from .algorithm import KernelIR
from .storage import DeviceView

__all__ = ["register", "configure", "routine_to_ir", "build_ir_library", "CudaModule", "BC_ROUTINES"]

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
        shift = sp.expand(idx - syms[0])           # (expand, not simplify: 19-27 populations x Q terms)
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
# module-like object returned to the driver
# ---------------------------------------------------------------------------
class CudaModule:
    """
    What `Generator.compile()` leaves in `generator.module` for backend "CUDA" (the extension module
    of autowrap.py:52-139): `.library` is the per-scheme kernel library (runtime.KernelLibrary), and
    every routine name is a callable following the kwargs protocol.
    `context` = {"dico", "algo", "storage", "compute", "lowering", "nv", "dim", "symmetric"} given by CudaSimulation;
    without it (a bare `autowrap(routines, 'cuda')`) the routines are lowered from their IR in fp64.
    """

    def __init__(self, routines, context=None, verbose=False, directory=None):
        context = context or {}
        self._context = context
        self._bc = {r.name: BC_ROUTINES[r.name] for r in routines if r.name in BC_ROUTINES}
        self._cell_routines = [r for r in routines if r.name not in BC_ROUTINES]
        self._device_lists = {}
        self.lowering = context.get("lowering") or "ir"
        self.source = None
        self.library = self.build(context.get("storage", "f64"), context.get("compute", "f64"),
                                  need_source=bool(verbose or directory))
        if verbose and self.source:
            print(self.source)
        if directory and self.source:
            import os

            os.makedirs(directory, exist_ok=True)
            with open(os.path.join(directory, os.path.basename(self.library.path)[3:-3] + ".cu"), "w") as fh:
                fh.write(self.source)
        if self.library is not None:
            for name in self.library.info["routines"]:
                setattr(self, name, self._kernel(name))
        for name, kind in self._bc.items():
            setattr(self, name, self._boundary(name, kind))

    # ---- lowering ----
    def _ir_kernels(self):
        kernels = [routine_to_ir(r) for r in self._cell_routines]
        by_name = {k.name: k for k in kernels}
        nconsm = self._context.get("nconsm")
        if "f2m" in by_name and nconsm and "f2m_consm" not in by_name:
            f2m = by_name["f2m"]         # conserved rows only: what `sol.m[symbol]` reads after a step
            kernels.append(KernelIR("f2m_consm", f2m.in_array, f2m.in_syms, f2m.in_offsets, f2m.out_array,
                                    f2m.statements, f2m.outputs[:nconsm], False, f2m.scalars))
        if "one_time_step" in by_name and self._context.get("symmetric") is not None:
            by_name["one_time_step"].symmetric = [int(k) for k in self._context["symmetric"]]
        return kernels

    def build(self, storage="f64", compute="f64", need_source=False):
        from . import build
        from .cudagen import generate_source

        if self.lowering == "scheme":
            from .simulation import build_kernel_library

            _, path, source = build_kernel_library(self._context["scheme"], self._context.get("settings"), storage,
                                                   need_source=need_source, compute=compute,
                                                   aa=bool(self._context.get("aa")))
            self.source = source or self.source
            return rt.KernelLibrary(path)
        kernels = self._ir_kernels()
        if not kernels:
            return None
        dim = len(kernels[0].in_offsets[0])
        nv = len(kernels[0].in_syms)
        c_type = {"f64": "double", "f32": "float"}
        source, info = generate_source(kernels, dim, nv, storage=c_type[storage], compute=c_type[compute])
        self.source = source
        return rt.KernelLibrary(build.build_kernels(source, info["hash"]))

    # ---- per-cell kernels: kwargs protocol of symbolic.py:288-299 ----
    def _kernel(self, name):
        lib = self.library
        info = lib.info["routines"][name]
        scalars = list(info["scalars"])
        in_name, out_name = info.get("in", "f"), info.get("out", "fnew")
        inner = bool(info.get("inner", False))

        def call(queue=None, **kw):
            src, dst = kw[in_name], kw[out_name]
            values = [float(kw[s]) for s in scalars]
            if isinstance(src, DeviceView):
                grid = src.dev.inner_grid() if inner else src.dev.grid
                rt.check(rt.lib().lbm_device_sync(), "sync")      # the steps run on the simulation's stream
                lib.launch(name, src.dev.ptr, dst.dev.ptr, grid, values)
                rt.check(rt.lib().lbm_device_sync(), "sync")
                return
            # small host arrays [nv, n, 1(, 1)] (wall equilibria: boundary.py:275-293)
            from .storage import DeviceArray

            if lib.info.get("storage", "double") != "double":
                raise NotImplementedError("host-array calls need the fp64 kernel library")
            ncell = int(np.prod(src.shape[1:]))
            dsrc = DeviceArray(src.shape[0], (ncell,), [0], "f64")
            dsrc.set(np.ascontiguousarray(src).reshape(src.shape[0], ncell))
            ddst = dsrc if dst is src else DeviceArray(dst.shape[0], (ncell,), [0], "f64")
            lib.launch(name, dsrc.ptr, ddst.ptr, dsrc.grid, values)
            rt.check(rt.lib().lbm_device_sync(), "sync")
            dst[...] = ddst.get().reshape(dst.shape)

        call.arg_dict = {k: None for k in dict.fromkeys([in_name, out_name] + scalars)}
        call.__name__ = name
        return call

    # ---- boundary kernels (compatibility protocol: `BoundaryMethod.update` of the reference calls
    #      them every step with host lists; CudaSimulation never does, its lists live on the device) ----
    def _cached(self, key, make):
        if key not in self._device_lists:
            host = make()
            ptr = ctypes.c_void_p()
            rt.check(rt.lib().lbm_malloc(ctypes.byref(ptr), max(8, host.nbytes)), "lbm_malloc")
            rt.check(rt.lib().lbm_memcpy_h2d(ptr, host.ctypes.data, host.nbytes), "h2d")
            self._device_lists[key] = (ptr.value, host.nbytes)
        return self._device_lists[key][0]

    def _positions(self, view, index):
        return self._cached(("pos", index.ctypes.data, index.shape),
                            lambda: np.ascontiguousarray(view.dev.positions(index.T)))

    def _values(self, tag, array, refresh):
        """fp64 list kept on the device; the buffer is reused, the content refreshed when asked."""
        array = np.ascontiguousarray(array, dtype=np.float64)
        key = (tag, array.ctypes.data, array.shape)
        fresh = key not in self._device_lists
        ptr = self._cached(key, lambda: array)
        if refresh and not fresh:
            rt.check(rt.lib().lbm_memcpy_h2d(ptr, array.ctypes.data, array.nbytes), "h2d")
        return ptr

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
            rhs = self._values("rhs", kw["rhs"], True) if kind != rt.BC_NEUMANN else None
            dist = self._values("dist", kw["dist"], False) if two_loads else None
            scratch = self._cached(("scratch", ncond), lambda: np.zeros(ncond))
            # the reference loop is sequential: gather-then-scatter keeps its result unless an entry
            # reads what an EARLIER entry of the same call stored (boundary.schedule handles that case
            # in CudaSimulation)
            rc = rt.lib().lbm_bc_apply(kind, f.dev.ptr, f.dev.storage_id, ncond, store, l0, l1, rhs, dist,
                                       scratch, 1, None)
            rt.check(rc, "lbm_bc_apply(%s)" % name)

        call.arg_dict = {k: None for k in names}
        call.__name__ = name
        return call

    def __del__(self):
        for ptr, _ in getattr(self, "_device_lists", {}).values():
            try:
                rt.lib().lbm_free(ptr)
            except Exception:
                pass


# ---------------------------------------------------------------------------
# process group discovery (the role of MPI.COMM_WORLD in the reference)
# ---------------------------------------------------------------------------
_config = {"slab": None, "nccl_id": None, "gather": None, "halo": "peer", "compute_dtype": None, "lowering": None}
_pending = []       # context of the CudaSimulation under construction, read by the Domain factory


def configure(**kwargs):
    """
    Process-wide settings of the CUDA backend (none is needed on one GPU):
      slab=(rank, nranks), nccl_id=bytes, gather=callable   explicit x-slab decomposition instead of the
                                                            torch.distributed discovery
      halo='peer' | 'nccl'                                  fused NVLink halo (default) or NCCL send/recv
      compute_dtype='float32'                               fp32 arithmetic (with dtype='float32')
      lowering='scheme' | 'ir'                              see the module docstring
    The last two can also be given per simulation as `dico['cuda_option'] = {'compute_dtype': ..,
    'lowering': .., 'in_place': True}` (the only key this backend adds to the reference's dictionary;
    `in_place` = in-place streaming, ONE population array).
    """
    for key, value in kwargs.items():
        if key not in _config:
            raise KeyError("unknown setting %r" % key)
        _config[key] = value


def _process_group():
    """(slab, nccl_id, gather) of this process."""
    if _config["slab"] is not None:
        return _config["slab"], _config["nccl_id"], (_config["gather"] if _config["halo"] == "peer" else None)
    import sys

    if "torch" not in sys.modules:
        # a process group can only exist if the caller imported torch: do not pay its import otherwise
        return None, None, None
    try:
        import torch.distributed as dist
    except Exception:
        return None, None, None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return None, None, None
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [None]
    if rank == 0:
        raw = (ctypes.c_char * 128)()
        rt.check(rt.lib().lbm_comm_unique_id(raw), "lbm_comm_unique_id")
        box[0] = bytes(raw.raw)
    dist.broadcast_object_list(box, src=0)

    def gather(blob):
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    return (rank, world), box[0], (gather if _config["halo"] == "peer" else None)


# ---------------------------------------------------------------------------
# registration
# ---------------------------------------------------------------------------
_registered = None


def _is_cuda(dico):
    return str(dico.get("generator", "")).upper() == "CUDA"


def register():
    """make `generator='cuda'` known to the importable pylbm (idempotent); returns the pylbm module."""
    global _registered
    if _registered is not None:
        return _registered
    import importlib

    import pylbm

    from . import domain as b200_domain
    from . import boundary as b200_boundary
    from .simulation import CudaContainer, CudaEngine

    # (pylbm.generator re-exports functions named like its sub-modules: go through importlib)
    ref_simulation = importlib.import_module("pylbm.simulation")
    ref_autowrap = importlib.import_module("pylbm.generator.autowrap")
    ref_codegen = importlib.import_module("pylbm.generator.codegen")
    ref_generator = importlib.import_module("pylbm.generator.generator")
    RefSimulation = ref_simulation.Simulation

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
            self.verbose, self.filepath = verbose, filepath

        def wrap_code(self, routines):
            return CudaModule(list(routines), None, self.verbose, self.filepath)

    original_wrapper = ref_autowrap.get_code_wrapper

    def get_code_wrapper(backend):
        if str(backend).upper() == "CUDA":
            return CudaCodeWrapper
        return original_wrapper(backend)

    ref_autowrap.get_code_wrapper = get_code_wrapper

    #    Generator.compile (generator.py:27-34) calls autowrap with the routines only; the CUDA module
    #    also wants to know the storage type and the scheme, which the Simulation left on the generator
    original_compile = ref_generator.Generator.compile

    def compile(self):
        if str(self.backend).upper() == "CUDA":
            self.module = CudaModule(list(self.routines.values()), getattr(self, "cuda_context", None),
                                     self.verbose, self.directory)
            return
        original_compile(self)

    ref_generator.Generator.compile = compile

    # 1. validator (validator.py:327-329): accept the new name
    try:
        ref_validator = importlib.import_module("pylbm.validator")
        original_validate = ref_validator.validate

        def validate(dico, name):
            if _is_cuda(dico):
                patched = dict(dico)
                patched["generator"] = "cython"
                patched.pop("cuda_option", None)
                return original_validate(patched, name)
            return original_validate(dico, name)

        ref_validator.validate = validate
        ref_simulation.validate = validate
    except Exception:       # pragma: no cover
        pass

    # 5. the two names the constructor resolves in its own module (simulation.py:94,136)
    RefDomain, RefBoundary = ref_simulation.Domain, ref_simulation.Boundary

    def Domain(dico, need_validation=True):
        if _is_cuda(dico):
            topology = _pending[-1].get("topology") if _pending else None
            return b200_domain.Domain(dico, need_validation=False, topology=topology,
                                      geometry_cls=pylbm.Geometry, stencil_cls=pylbm.Stencil)
        return RefDomain(dico, need_validation=need_validation)

    def Boundary(domain, generator, dico):
        if str(getattr(generator, "backend", "")).upper() == "CUDA":
            return b200_boundary.Boundary(domain, generator, dico)
        return RefBoundary(domain, generator, dico)

    ref_simulation.Domain = Domain
    ref_simulation.Boundary = Boundary

    class CudaSimulation(CudaEngine, RefSimulation):
        """`pylbm.Simulation(dico)` for generator='cuda' (see the module docstring)."""

        def __init__(self, dico, sorder=None, dtype="float64", check_inverse=False, initialize=True):
            rt.ensure_gpu()
            option = dico.get("cuda_option") or {}
            storage, compute = self._storage_names(dtype, option.get("compute_dtype", _config["compute_dtype"]))
            slab, nccl_id, gather = _process_group()
            self._engine_defaults(storage, compute, slab, nccl_id, gather, bool(option.get("in_place", False)))
            self._lowering = option.get("lowering", _config["lowering"])
            topology = None
            if self.nranks > 1:
                topology = b200_domain.SlabTopology(pylbm.Stencil.extract_dim(dico), self.rank, self.nranks)
                topology.gather = gather
            _pending.append({"topology": topology})
            try:
                RefSimulation.__init__(self, dico, sorder, dtype, check_inverse, initialize)
            finally:
                _pending.pop()

        # container table (simulation.py:165-171)
        def _get_container(self, sorder):
            return CudaContainer(self.domain, self.scheme, sorder, self.storage, self.in_place)

        # algorithm (simulation.py:179-190): unchanged, but the generator learns what it compiles for
        def _get_algorithm(self, dico, sorder):
            algo = RefSimulation._get_algorithm(self, dico, sorder)
            stock = type(algo) is importlib.import_module("pylbm.algorithm").PullAlgorithm
            lowering = self._lowering or ("scheme" if stock else "ir")
            if lowering not in ("scheme", "ir"):
                raise ValueError("cuda_option['lowering'] must be 'scheme' or 'ir', got %r" % (lowering,))
            if lowering == "scheme" and not stock:
                raise ValueError("cuda_option['lowering']='scheme' only knows the stock PullAlgorithm")
            if self.in_place and lowering != "scheme":
                raise NotImplementedError("in-place streaming is generated by the scheme lowering only")
            import os

            context = {"storage": self.storage, "compute": self.compute, "lowering": lowering,
                       "aa": self.in_place or bool(os.environ.get("PYLBM_B200_AA_LIBRARY")),
                       "nconsm": len(self.scheme.consm), "symmetric": self.scheme.stencil.get_symmetric(),
                       "settings": dict(algo.settings) if hasattr(algo, "settings") else None}
            if lowering == "scheme":
                context["scheme"] = _scheme_twin(dico, self.scheme)
                # the scheme lowering does not consume the reference's symbolic routines: building them
                # (algorithm/base.py:602-628, several seconds of sympy for a 3-D scheme) is deferred until
                # somebody asks for them (`sol.algo.generate()` still works, `lowering='ir'` runs it)
                # (the closure must not capture the simulation: a reference cycle would keep its device
                # arrays alive until the next garbage collection)
                reference_generate, generator = algo.generate, self.generator

                def generate_routines():
                    if not generator.routines:
                        reference_generate()

                algo.generate = lambda: None
                algo.generate_routines = generate_routines
            self.generator.cuda_context = context
            return algo

        kernels = property(lambda self: self.generator.module.library)

        def _build_kernels(self, storage, compute):
            return self.generator.module.build(storage, compute)

        def __repr__(self):
            return RefSimulation.__str__(self)

        __str__ = __repr__

    CudaSimulation.__module__ = __name__
    CudaSimulation.__qualname__ = "CudaSimulation"

    def __new__(cls, dico=None, *args, **kwargs):
        if cls is RefSimulation and isinstance(dico, dict) and _is_cuda(dico):
            return object.__new__(CudaSimulation)
        return object.__new__(cls)

    RefSimulation.__new__ = staticmethod(__new__)
    globals()["CudaSimulation"] = CudaSimulation
    _registered = pylbm
    return pylbm


def build_ir_library(dico, storage="f64", compute="f64"):
    """
    The kernel library of a dictionary lowered from the reference's own Routines, without touching a
    device: what `pylbm.Simulation(dico)` with `cuda_option={'lowering': 'ir'}` compiles.  Used by the
    build box (`__graft_entry__.build()`) to fill the in-tree cache.  Returns the CudaModule.
    """
    import importlib

    pylbm = register()
    generator_cls = importlib.import_module("pylbm.generator.generator").Generator
    algo_cls = importlib.import_module("pylbm.algorithm").PullAlgorithm
    scheme = pylbm.Scheme(dico)
    generator = generator_cls("CUDA")
    algo_cls(scheme, list(range(scheme.dim + 1)), generator,
             {"m_local": True, "split": False, "check_isfluid": False}).generate()
    context = {"storage": storage, "compute": compute, "lowering": "ir", "nconsm": len(scheme.consm),
               "symmetric": scheme.stencil.get_symmetric()}
    return CudaModule(list(generator.routines.values()), context)


def _scheme_twin(dico, ref_scheme):
    """
    Numeric twin of the reference Scheme for the scheme-level lowering (exact rational M and inverse,
    parameters substituted), built from the same dictionary and cross-checked against the reference's
    object: same populations, same conserved moments at the same rows, same moment matrix.
    """
    from .scheme import Scheme

    twin = Scheme(dico, need_validation=False)
    ref_consm = {str(k): int(v) for k, v in ref_scheme.consm.items()}
    if {str(k): int(v) for k, v in twin.consm.items()} != ref_consm:
        raise rt.LbmError("scheme lowering: conserved moments differ from the reference's (%s)" % ref_consm)
    if int(twin.stencil.nv_ptr[-1]) != int(ref_scheme.stencil.nv_ptr[-1]):
        raise rt.LbmError("scheme lowering: number of populations differs from the reference's")
    ref_M = sp.Matrix(ref_scheme.M).subs(list(ref_scheme.param.items()))
    M = np.array(ref_M.evalf(), dtype=np.float64)
    mine = np.array(sp.Matrix(twin.M).evalf(), dtype=np.float64)
    if M.shape != mine.shape or not np.allclose(M, mine, rtol=1e-13, atol=1e-13):
        raise rt.LbmError("scheme lowering: moment matrix differs from the reference's")
    return twin
