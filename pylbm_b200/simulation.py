"""
Simulation driver with the reference's public surface, running on one B200 per
process through the C ABI of include/lbm_b200.h.

Mirror of pylbm.simulation.Simulation (reference: pylbm/simulation.py:89-153
construction, 155-163 _initialize, 200-243 m/F item properties, 258-320
initialization, 322-371 kernel wrappers, 373-390 boundary_condition, 392-420
one_time_step).  Same dictionary, same attribute names, same order of
operations per time step; what changes is where the work happens:

* the kernels come from a generated CUDA library (cudagen.py) instead of a
  generated Cython module; F, Fnew and m stay in HBM (storage.DeviceArray);
* `one_time_step()` is a single enqueue-only call into the runtime
  (`lbm_sim_step`): ghost update, boundary kernels, fused pull kernel, swap.
  Constant right-hand sides are uploaded once instead of being recomputed with
  NumPy every step (reference: boundary.py:421-427 runs every step);
* `run(nsteps)` enqueues many steps in one call (CUDA-graph pairs when the
  kernels do not depend on t).
"""

import ctypes
import os
import types

import numpy as np
import sympy as sp

from . import runtime as rt
from . import build
from .algorithm import PullAlgorithm
from .boundary import Boundary
from .cudagen import generate_source, kernel_tag
from .domain import Domain, SlabTopology
from .scheme import Scheme
from .storage import DeviceArray

__all__ = ["Simulation", "CudaEngine", "CudaContainer"]


def build_kernel_library(scheme, settings=None, storage="f64", need_source=False, compute="f64", aa=False):
    """
    scheme -> per-cell kernel IR -> CUDA C -> liblbmk_<hash>.so (cached in-tree by source hash).
    Needs nvcc but no GPU, so it is also what `__graft_entry__.build()` runs on the build box.
    `aa`: library for in-place streaming (even / odd step launcher, swapped-read moment kernels).
    Returns (algorithm, library path, CUDA source).
    """
    algo_settings = {"m_local": True, "split": False, "check_isfluid": False}
    algo_settings.update(settings or {})
    algo = PullAlgorithm(scheme, algo_settings)
    c_storage = "double" if storage == "f64" else "float"
    c_compute = "double" if compute == "f64" else "float"
    kernels = algo.kernels(aa=aa)
    path = build.kernel_library_path(kernel_tag(kernels, scheme.dim, algo.ns, storage=c_storage, compute=c_compute,
                                                aa=aa))
    if os.path.exists(path) and not need_source:
        return algo, path, None          # cached: no lowering, no nvcc
    source, info = generate_source(kernels, scheme.dim, algo.ns, storage=c_storage, compute=c_compute, aa=aa)
    return algo, build.build_kernels(source, info["hash"]), source


class _ItemProperty:
    """`sol.m[key]` / `sol.m[key] = value` style access (reference: utils.py:26-80)."""

    def __init__(self, owner, getter, setter=None):
        self._owner, self._getter, self._setter = owner, getter, setter

    def __getitem__(self, key):
        return self._getter(self._owner, key)

    def __setitem__(self, key, value):
        if self._setter is None:
            raise AttributeError("read-only item property")
        self._setter(self._owner, key, value)


class CudaContainer:
    """m, F, Fnew in HBM (the 'cuda' entry next to the reference's Numpy/Cython/Loopy
    containers: pylbm/container.py:10-106)."""

    gpu_support = True

    def __init__(self, domain, scheme, sorder=None, storage="f64", in_place=False):
        self.dim = domain.dim
        self.mpi_topo = domain.mpi_topo
        self.nv = int(scheme.stencil.nv_ptr[-1])
        self.nspace = domain.global_size
        self.vmax = list(domain.stencil.vmax)
        if sorder is not None and sorted(int(i) for i in sorder) != list(range(self.dim + 1)):
            raise ValueError("sorder must be a permutation of range(dim + 1), got %r" % (sorder,))
        # one device layout ([population][x][y][z], padded rows): the reference's `sorder` permutes a
        # host array (storage.py:60-118); here the argument is validated and has no effect
        self.sorder = [i for i in range(self.dim + 1)]
        shape = domain.shape_halo
        self.F = DeviceArray(self.nv, shape, self.vmax, storage, scheme.consm)
        # in-place streaming (AA pattern): ONE population array
        self.Fnew = self.F if in_place else DeviceArray(self.nv, shape, self.vmax, storage, scheme.consm)
        self.in_place = bool(in_place)
        self._m = None
        self._storage = storage
        self._shape = shape
        self._consm = scheme.consm

    @property
    def m(self):
        # moments are only materialised when somebody asks for them
        if self._m is None:
            # same element layout as F (one lbmk_grid addresses both in f2m / m2f)
            self._m = DeviceArray(self.nv, self._shape, self.vmax, "f64", self._consm, align=self.F.align)
        return self._m

    @property
    def mc(self):
        """conserved moments only (nconsm rows, same element layout): target of `f2m_consm`."""
        if getattr(self, "_mc", None) is None:
            self._mc = DeviceArray(len(self._consm), self._shape, self.vmax, "f64", self._consm, align=self.F.align)
        return self._mc

    def move2gpu(self, array):
        """host arrays of the front end stay where they are: what the kernels need of them (boundary
        lists) is uploaded once by BoundaryMethod.move2gpu (reference: container.py:100-106)."""
        return array

    def release_m(self):
        """free the moment array (it is rebuilt by f2m on the next `sol.m[...]`)."""
        self._m = None


class CudaEngine:
    """
    The device side of a simulation: runtime handle, boundary lists on the device, kernel launches,
    item properties, time stepping.  It is a mixin over an object that owns `domain`, `scheme`,
    `container` (CudaContainer), `bc` (Boundary of this package), `kernels` (runtime.KernelLibrary),
    `init_type` / `init_data` and the scalars t, nt, dt_, dim, extra_parameters -- i.e. what the
    reference's constructor builds (pylbm/simulation.py:89-153).  Two front ends use it:
    `pylbm_b200.Simulation` (this package's own dictionary front end, no pylbm needed) and
    `pylbm_b200.plugin.CudaSimulation`, the class `pylbm.Simulation(dico)` instantiates for
    generator='cuda' once `plugin.register()` ran (the reference's own constructor, unchanged).
    """

    @staticmethod
    def _storage_names(dtype, compute_dtype):
        names = {"float64": "f64", "f64": "f64", "float32": "f32", "f32": "f32"}
        try:
            storage = names[dtype if isinstance(dtype, str) else str(np.dtype(dtype))]
        except KeyError:
            raise ValueError("dtype must be float64 or float32 (storage of the populations), got %r" % (dtype,))
        try:
            compute = "f64" if compute_dtype is None else names[
                compute_dtype if isinstance(compute_dtype, str) else str(np.dtype(compute_dtype))]
        except KeyError:
            raise ValueError("compute_dtype must be float64 or float32, got %r" % (compute_dtype,))
        if compute == "f32" and storage != "f32":
            raise ValueError("compute_dtype='float32' needs dtype='float32' (fp32 storage of the populations)")
        return storage, compute

    def _engine_defaults(self, storage, compute, slab=None, nccl_id=None, gather=None, in_place=False):
        self.storage, self.compute = storage, compute
        self.in_place = bool(in_place)
        if self.in_place and slab is not None and slab[1] > 1 and gather is not None:
            raise ValueError("in-place streaming on several GPUs uses the NCCL halo (no `gather` / halo='nccl'): "
                             "the fused NVLink halo stores into the neighbours' second array")
        self.rank, self.nranks = slab if slab is not None else (0, 1)
        self._nccl_id, self._gather = nccl_id, gather
        self._mc_version = -1
        self._handle = None
        self._time_dependent = False

    # ------------------------------------------------------------------
    # `_update_m = True` means "F changed, the moments are stale" (reference: simulation.py:215-224);
    # every such assignment also invalidates the conserved-only copy
    @property
    def _update_m(self):
        return self.__dict__.get("_update_m_flag", True)

    @_update_m.setter
    def _update_m(self, value):
        self.__dict__["_update_m_flag"] = bool(value)
        if value:
            self.__dict__["_f_version"] = self.__dict__.get("_f_version", 0) + 1

    @property
    def _f_version(self):
        return self.__dict__.get("_f_version", 0)

    @property
    def dt(self):
        if isinstance(self.dt_, sp.Expr):
            subs = list(self.scheme.param.items()) + list(self.extra_parameters.items())
            self.dt_ = float(self.dt_.subs(subs))
        return self.dt_

    def _scalar_values(self, name):
        out = []
        extra = {str(k): v for k, v in self.extra_parameters.items()}
        for s in self.kernels.scalars(name):
            if s == "t":
                out.append(self.t)
            elif s == "dt":
                out.append(self.dt)
            elif s in extra:
                out.append(float(extra[s]))
            else:
                raise KeyError(
                    "the kernel %s needs a value for the symbol %r: give it in 'parameters' or in "
                    "sol.extra_parameters" % (name, s)
                )
        return out

    # ---- runtime object -------------------------------------------------
    def _create_handle(self):
        F, Fnew = self.container.F, self.container.Fnew
        desc = rt.LbmSimDesc()
        desc.nv = F.nv
        desc.storage = F.storage_id
        desc.grid = F.inner_grid()
        for a in range(3):
            desc.vmax[a] = F.canonical_vmax[a]
        desc.periodic_mask = sum(1 << a for a in range(3) if F.canonical_vmax[a] > 0)
        desc.f, desc.fnew = F.ptr, Fnew.ptr
        desc.one_time_step = self.kernels.address("one_time_step")
        desc.one_time_step_peers = self.kernels.address("one_time_step_peers")
        names = self.kernels.scalars("one_time_step")
        desc.nscalars = len(names)
        desc.t_index = names.index("t") if "t" in names else -1
        try:
            values = self._scalar_values("one_time_step")
            self._scalars_unresolved = False
        except KeyError:
            # user parameters may be given later (extra_parameters): resolved, loudly, by the first step
            values = [0.0] * len(names)
            self._scalars_unresolved = True
        for i, v in enumerate(values):
            desc.scalars[i] = v
        desc.t, desc.dt = self.t, self.dt
        vel = self.scheme.stencil.get_all_velocities()
        for k in range(F.nv):
            for d in range(self.dim):
                desc.vel[k][3 - self.dim + d] = int(vel[k][d])
        handle = rt.lib().lbm_sim_create(ctypes.byref(desc))
        if not handle:
            raise rt.LbmError("lbm_sim_create failed: %s" % rt.lib().lbm_last_error().decode())
        self._handle = handle
        if self.nranks > 1:
            rt.check(rt.lib().lbm_sim_comm_init(handle, self.rank, self.nranks, self._nccl_id), "lbm_sim_comm_init")
            if self._gather is not None:
                # direct NVLink halo: swap CUDA-IPC handles with the two ring neighbours
                blob = (ctypes.c_char * 256)()
                rt.check(rt.lib().lbm_sim_ipc_export(handle, blob), "lbm_sim_ipc_export")
                blobs = self._gather(bytes(blob.raw))      # list of the blobs of all ranks, by rank
                left = blobs[(self.rank - 1) % self.nranks]
                right = blobs[(self.rank + 1) % self.nranks]
                rt.check(rt.lib().lbm_sim_ipc_open(handle, left, right), "lbm_sim_ipc_open")
        self._time_dependent = False

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle:
            try:
                rt.lib().lbm_sim_destroy(handle)
            except Exception:
                pass
            self._handle = None

    def _initialize(self):
        """(reference: simulation.py:155-163)"""
        self._create_handle()
        self.initialization()
        F = self.container.F
        for method in self.bc.methods:
            method.prepare_rhs(self)
            method.fix_iload()
            method.set_rhs()
            method.prepare_device(F)
        tasks = None if self.in_place else self._plan_tasks()
        masks = self._plan_walls() if tasks is None else [None] * len(self.bc.methods)
        # device methods in application order; the entries a wall plan takes out of a method follow it
        # immediately as a stale-only method (same place in the sequence as in the reference)
        info = []
        self._wall_keep = []
        lib = rt.lib()
        for method, mask in zip(self.bc.methods, masks):
            method.move2gpu(self._handle, F, wall_mask=mask)
            store, l0, l1, _, _, level_ptr, two_phase = method._keep
            single = len(level_ptr) == 2 and not int(two_phase[0])
            info.append((store, [l0] + ([l1] if l1 is not None else []), single))
            if mask is not None and mask.any():
                wstore, wl0, wrhs = method._wall_lists
                idx = lib.lbm_sim_add_bc(self._handle, method.kind, wstore.size, wstore.ctypes.data, wl0.ctypes.data,
                                         None, wrhs.ctypes.data, None, 0, None, None)
                rt.check(idx, "lbm_sim_add_bc(wall entries)")
                rt.check(lib.lbm_sim_bc_stale_only(self._handle, idx, 1), "lbm_sim_bc_stale_only")
                self._wall_keep.append((wstore, wl0, wrhs, idx))
                info.append((wstore, [wl0], False))          # never merged with its neighbours
        walls_desc = None
        if self.bc.walls is not None:
            w = walls_desc = rt.LbmkWalls()
            w.lo_plane, w.hi_plane = self.bc.walls["lo_plane"], self.bc.walls["hi_plane"]
            w.neg_lo, w.neg_hi = self.bc.walls["neg_lo"], self.bc.walls["neg_hi"]
            for k in range(64):
                w.rhs[k] = float(self.bc.walls["rhs"][k])
            if not self.in_place:
                rt.check(lib.lbm_sim_set_walls(self._handle, self.kernels.address("one_time_step_walls"),
                                               ctypes.byref(w)), "lbm_sim_set_walls")
        if len(info) > 1:
            # consecutive methods that provably do not interact run as one kernel launch
            from .boundary import merge_groups

            groups = merge_groups(info)
            if len(groups) - 1 < len(info):
                rt.check(lib.lbm_sim_bc_groups(self._handle, len(groups) - 1,
                                               groups.ctypes.data_as(ctypes.POINTER(ctypes.c_int))),
                         "lbm_sim_bc_groups")
            self.bc.groups = groups
        if tasks is not None:
            ptr = lambda a: a.ctypes.data
            rt.check(lib.lbm_sim_set_tasks(
                self._handle, self.kernels.address("one_time_step_tasks"), tasks["ntasks"], tasks["nblocks"],
                ptr(tasks["block_ptr"]), ptr(tasks["code"]), ptr(tasks["l0"]), ptr(tasks["l1"]), ptr(tasks["dist"]),
                ptr(tasks["ibc"]), ptr(tasks["entry"]), tasks["ngroups_y"], tasks["ngroups_x"], tasks["tx"]),
                "lbm_sim_set_tasks")
        self.bc.tasks = None if tasks is None else {k: tasks[k] for k in ("ntasks", "nentries", "nblocks")}
        if self.in_place:
            self._enable_in_place(walls_desc)
        if not os.environ.get("PYLBM_B200_HOST_TIME_BC"):
            for method in self.bc.methods:
                method.prepare_time_bc(self)
        self._time_dependent = any(m.is_time_dependent for m in self.bc.methods)
        self._need_init = False

    def _enable_in_place(self, walls_desc=None):
        """in-place streaming: the lists of the odd steps (boundary.plan_aa) for every registered method --
        the stale-only wall entries included -- and the even / odd launcher (with the fused walls when a
        wall plan was accepted)."""
        from .boundary import plan_aa

        lib = rt.lib()
        info, index = [], []
        for method in self.bc.methods:
            store, l0, l1, _, _, _, _ = method._keep
            info.append({"store": store, "loads": [l0] + ([l1] if l1 is not None else [])})
            index.append(method.device_index)
        for wstore, wl0, _, idx in self._wall_keep:
            info.append({"store": wstore, "loads": [wl0]})
            index.append(idx)
        odd = plan_aa(info, self.container.F, self.scheme.stencil.get_all_velocities(),
                      self.scheme.stencil.get_symmetric())
        if odd is None:
            raise NotImplementedError("in-place streaming: a boundary entry touches the outermost ghost layer "
                                      "against its own velocity")
        self._odd_keep = []
        for idx, lists in zip(index, odd):
            st = np.ascontiguousarray(lists["store"], dtype=np.int64)
            ld = [np.ascontiguousarray(l, dtype=np.int64) for l in lists["loads"]]
            self._odd_keep.append((st, ld))
            rt.check(lib.lbm_sim_set_bc_odd(self._handle, idx, st.ctypes.data, ld[0].ctypes.data,
                                            ld[1].ctypes.data if len(ld) > 1 else None), "lbm_sim_set_bc_odd")
        rt.check(lib.lbm_sim_set_aa(self._handle, self.kernels.address("one_time_step_aa")), "lbm_sim_set_aa")
        if walls_desc is not None:
            rt.check(lib.lbm_sim_set_aa_walls(self._handle, self.kernels.address("one_time_step_aa_walls"),
                                              ctypes.byref(walls_desc)), "lbm_sim_set_aa_walls")

    @property
    def _swapped(self):
        """the in-place array is in the swapped layout (an odd number of steps since the last even one)."""
        return bool(self.in_place and self._handle and rt.lib().lbm_sim_aa_phase(self._handle) == 1)

    def _need_natural(self, what):
        if self._swapped:
            raise RuntimeError("in-place streaming: %s needs the natural layout of the populations, i.e. an even "
                               "number of time steps" % what)

    def _plan_tasks(self):
        """boundary entries evaluated by the fused kernel itself (boundary.plan_tasks): one launch per
        step instead of list kernels + fused kernel.  OPT-IN (PYLBM_B200_TASKS=1): measured on B200 it
        loses everywhere -- the entries of a block need three dependent memory round trips (task range,
        task records, populations) before the block can start its collision, which costs more than the
        list kernels it removes (D2Q9 256^2: 7.2 vs 5.1 us/step; Karman 4096x1024: 103.0 vs 100.5 us;
        D3Q19 256^3: 1.10 vs 0.82 ms; profiles/r02_tasks_vs_lists.md).  Kept because it is bit-identical
        and tested, and because it makes a step a single launch (useful under a debugger / ncu)."""
        from .boundary import plan_tasks

        self.bc.walls = None
        if os.environ.get("PYLBM_B200_TASKS", "0") != "1" or not self.bc.methods:
            return None
        F = self.container.F
        info = []
        for method in self.bc.methods:
            store, l0, l1, _, dist, level_ptr, two_phase = method._keep
            info.append({"kind": method.kind, "store": store, "loads": [l0] + ([l1] if l1 is not None else []),
                         "dist": dist, "single": len(level_ptr) == 2 and not int(two_phase[0])})
        return plan_tasks(info, F, self.scheme.stencil.get_all_velocities())

    def _plan_walls(self):
        """bounce-back walls of the fastest axis applied by the fused kernel (boundary.plan_walls)."""
        from .boundary import plan_walls

        self.bc.walls = None
        none = [None] * len(self.bc.methods)
        if os.environ.get("PYLBM_B200_NO_WALLS") or not self.bc.methods:
            return none
        if self.nranks > 1 and (self.dim == 1 or self._gather is None):
            # 1-D: the fastest axis is the slab axis.  NCCL halo: whole slab-face planes are copied
            # at the start of a step, ghost rows of the fastest axis included, which would overwrite the
            # corner values the kernel stored at the end of the previous step (the peer halo only
            # stores the populations that cross the slab face)
            return none
        info = []
        for method in self.bc.methods:
            store, l0, l1, rhs, _, level_ptr, two_phase = method._keep
            eligible = len(level_ptr) == 2 and not int(two_phase[0]) and not method.is_time_dependent
            info.append({"kind": method.kind, "store": store, "loads": [l0] + ([l1] if l1 is not None else []),
                         "rhs": rhs, "eligible": eligible})
        plan = plan_walls(info, self.container.F, self.scheme.stencil.get_all_velocities(),
                          self.scheme.stencil.get_symmetric())
        if plan is None:
            return none
        self.bc.walls, masks = plan
        return masks

    def initialization(self):
        """(reference: simulation.py:258-320)"""
        coords = np.meshgrid(*(c for c in self.domain.coords_halo), sparse=True, indexing="ij")
        if self.init_type == "moments":
            target = self.container.m
        elif self.init_type == "distributions":
            target = self.container.F
        else:
            raise ValueError("the key `inittype` should be moments or distributions")
        if self.init_data is None:
            return
        for k, v in self.init_data.items():
            if isinstance(v, tuple):
                f = v[0]
                extraargs = v[1] if len(v) == 2 else ()
                target[k] = f(*(tuple(coords) + tuple(extraargs)))
            elif isinstance(v, types.FunctionType):
                target[k] = v(*coords)
            else:
                target[k] = v
        if self.init_type == "moments":
            self.equilibrium()
            self.m2f()
        else:
            self.f2m()
        if self.container.Fnew is not self.container.F:
            self.container.Fnew.copy_from(self.container.F)
        self._invalidate_ghosts()
        self._update_m = True
        self.container.release_m()   # rebuilt on demand by f2m; frees nv * cells * 8 bytes of HBM

    # ---- whole-array kernels (reference: simulation.py:322-371) -------------
    def _launch(self, name, src, dst, inner=False, kernels=None):
        grid = src.inner_grid() if inner else src.grid
        stream = rt.lib().lbm_sim_stream(self._handle) if self._handle else None
        (kernels or self.kernels).launch(name, src.ptr, dst.ptr, grid, self._scalar_values(name), stream)
        if self._handle:
            rt.check(rt.lib().lbm_sim_sync(self._handle), "sync")
        else:
            rt.check(rt.lib().lbm_device_sync(), "sync")

    def _kernels_f64(self):
        if self.storage == "f64":
            return self.kernels
        if self.__dict__.get("_kernels64") is None:
            self._kernels64 = self._build_kernels("f64", "f64")
        return self._kernels64

    def _scratch(self, nv, ncell, storage, slot):
        """cached 1-D device arrays for the small wall-value evaluations (time-dependent boundary
        values call them every step)."""
        cache = self.__dict__.setdefault("_scratch_arrays", {})
        key = (nv, ncell, storage, slot)
        if key not in cache:
            cache[key] = DeviceArray(nv, (ncell,), [0], storage, align=self.container.F.align)
        return cache[key]

    def _on_device(self, host):
        """1-D device twin of a small HostArray (kernels are shape agnostic)."""
        ncell = int(np.prod(host.nspace))
        dev = self._scratch(host.nv, ncell, "f64", 0)
        dev.set(host.array.reshape(host.nv, ncell))
        return dev

    def _invalidate_ghosts(self):
        """F was written from outside the step loop: the next step must refresh its ghost layers."""
        if self._handle:
            rt.check(rt.lib().lbm_sim_invalidate_ghosts(self._handle), "lbm_sim_invalidate_ghosts")

    def f2m(self, **kwargs):
        if self._swapped:      # interior cells, read through the swapped layout
            self._launch("f2m_sw", self.container.F, self.container.m, inner=True)
            return
        self._launch("f2m", self.container.F, self.container.m)

    def m2f(self, m_user=None, f_user=None, **kwargs):
        if m_user is not None:
            dm = self._on_device(m_user)
            # wall equilibria stay fp64: with fp32 populations the m2f of the fp64-storage library of the
            # same scheme is used (rhs = feq[k] -/+ feq[ksym] cancels, boundary.py:421-427)
            df = self._scratch(dm.nv, dm.nspace[0], "f64", 1)
            self._launch("m2f", dm, df, kernels=self._kernels_f64())
            f_user.array[...] = df.get().reshape(f_user.array.shape)
            return
        self._need_natural("m2f")
        self._launch("m2f", self.container.m, self.container.F)
        self._invalidate_ghosts()

    def equilibrium(self, m_user=None, **kwargs):
        if m_user is not None:
            dm = self._on_device(m_user)
            self._launch("equilibrium", dm, dm)
            m_user.array[...] = dm.get().reshape(m_user.array.shape)
            return
        self._launch("equilibrium", self.container.m, self.container.m)

    def relaxation(self, **kwargs):
        self._launch("relaxation", self.container.m, self.container.m)

    def source_term(self, fraction_of_time_step=1.0, **kwargs):
        """half a time step of explicit Euler on the moments (ode.py:11-16: `m + dt/2 * rhs`), whatever
        `fraction_of_time_step` says -- the reference does not forward that argument to its kernel
        either (simulation.py:340-345)."""
        if "source_term" not in self.kernels.routines:
            raise KeyError("source_term: the scheme has no source term")
        self._launch("source_term", self.container.m, self.container.m, inner=True)

    def transport(self, **kwargs):
        """Fnew <- streamed F, then F <- Fnew: the result is left in F ("the array _F is modified",
        simulation.py:322-327, and what the reference's NumPy backend does; its Cython backend leaves
        the result in Fnew, base.py:289-296)."""
        F, Fnew = self.container.F, self.container.Fnew
        if Fnew is F:
            raise NotImplementedError("the stand-alone transport needs two arrays (in_place=False)")
        self._launch("transport", F, Fnew, inner=True)
        F.copy_from(Fnew)
        self._invalidate_ghosts()

    # ---- item properties (reference: simulation.py:200-243) -----------------
    def _refresh_m(self):
        if self._update_m:
            self._update_m = False
            self.f2m()

    def _conserved(self, key):
        """interior values of one conserved moment, computed from F by the conserved-only kernel
        (nconsm instead of Q rows of device memory and of store traffic)."""
        c = self.container
        if self._mc_version != self._f_version:
            if self._swapped:
                self._launch("f2m_consm_sw", c.F, c.mc, inner=True)
            else:
                self._launch("f2m_consm", c.F, c.mc)
            self._mc_version = self._f_version
        return c.mc._in(key)

    @property
    def m_halo(self):
        def get(self_, i):
            self_._refresh_m()
            return self_.container.m[i]

        def put(self_, i, value):
            # the moment array is rebuilt lazily (release_m): bring it up to date before one row is
            # overwritten, so that the other rows keep the moments of the last f2m like the reference's
            if self_.container._m is None or self_._update_m:
                self_._update_m = False
                self_.f2m()
            self_._update_m = False
            self_.container.m[i] = value

        return _ItemProperty(self, get, put)

    @property
    def m(self):
        def get(self_, i):
            row = self_.container.F._key(i)
            try:
                row = int(row) if not isinstance(row, (slice, tuple, list, np.ndarray)) else -1
            except (TypeError, ValueError):
                row = -1
            if self_._update_m and 0 <= row < len(self_.scheme.consm):
                return self_._conserved(row)
            self_._refresh_m()
            return self_.container.m._in(i)

        return _ItemProperty(self, get)

    @property
    def F_halo(self):
        def get(self_, i):
            self_._need_natural("F_halo")
            return self_.container.F[i]

        def put(self_, i, value):
            self_._need_natural("writing F")
            self_._update_m = True
            self_.container.F[i] = value
            self_._invalidate_ghosts()

        return _ItemProperty(self, get, put)

    @property
    def F(self):
        def get(self_, i):
            if not self_._swapped:
                return self_.container.F._in(i)
            # swapped in-place array: population k of cell x sits in slot (kbar, x + v_k)
            F = self_.container.F
            k = int(F._key(i))
            sym = self_.scheme.stencil.get_symmetric()
            v = self_.scheme.stencil.get_all_velocities()[k]
            whole = F.get(int(sym[k]), 1)[0]
            sl = tuple(slice(w + int(v[d]), whole.shape[d] - w + int(v[d])) for d, w in enumerate(F.vmax))
            return whole[sl].copy()

        return _ItemProperty(self, get)

    # ---- time stepping ------------------------------------------------------
    def _push_scalars(self):
        names = self.kernels.scalars("one_time_step")
        if names and names != ["t"]:
            values = self._scalar_values("one_time_step")      # KeyError for a symbol without a value
            arr = (ctypes.c_double * len(values))(*values)
            rt.check(rt.lib().lbm_sim_set_scalars(self._handle, arr, len(values)), "lbm_sim_set_scalars")
        self._scalars_unresolved = False

    def _update_time_bc(self):
        for method in self.bc.methods:
            if not method.is_time_dependent:
                continue
            if getattr(method, "_time_plans", None):
                method.update_feq_device(self)      # only the user's callback runs on the host
            else:
                method.update_feq(self)
                method.set_rhs()
                method.push_rhs(self._handle)

    def boundary_condition(self, **kwargs):
        """ghost update + boundary methods on F (reference: simulation.py:373-390)."""
        if self._need_init:
            self._initialize()
        self._update_time_bc()
        rt.check(rt.lib().lbm_sim_boundary_condition(self._handle), "lbm_sim_boundary_condition")

    def one_time_step(self, **kwargs):
        """(reference: simulation.py:392-420)"""
        if self._need_init:
            self._initialize()
        self._update_m = True
        if self._time_dependent:
            self._update_time_bc()
        if self.extra_parameters or self._scalars_unresolved:
            self._push_scalars()
        rc = rt.lib().lbm_sim_step(self._handle, 1)
        if rc < 0:
            rt.check(rc, "lbm_sim_step")
        c = self.container
        c.F, c.Fnew = c.Fnew, c.F
        self.t += self.dt
        self.nt += 1

    def run(self, nsteps, graph=True):
        """nsteps time steps enqueued by ONE runtime call (falls back to a Python loop when a
        boundary value depends on time)."""
        if self._need_init:
            self._initialize()
        if self._time_dependent:
            for _ in range(nsteps):
                self.one_time_step()
            return
        self._update_m = True
        if self.extra_parameters or self._scalars_unresolved:
            self._push_scalars()
        rt.check(rt.lib().lbm_sim_use_graph(self._handle, 1 if graph else 0), "lbm_sim_use_graph")
        rt.check(rt.lib().lbm_sim_step(self._handle, int(nsteps)), "lbm_sim_step")
        if nsteps % 2:
            c = self.container
            c.F, c.Fnew = c.Fnew, c.F
        for _ in range(nsteps):
            self.t += self.dt
        self.nt += nsteps

    def synchronize(self):
        if self._handle:
            rt.check(rt.lib().lbm_sim_sync(self._handle), "lbm_sim_sync")

    def __repr__(self):
        return "Simulation(generator='cuda', {}, {}, t={}, nt={})".format(self.domain, self.scheme, self.t, self.nt)


class Simulation(CudaEngine):
    """
    Simulation(dico, sorder=None, dtype='float64', check_inverse=False, initialize=True,
               slab=None, nccl_id=None, gather=None, compute_dtype=None)

    `dico` is a pylbm dictionary (box, elements, space_step, scheme_velocity, schemes,
    parameters, relative_velocity, init, inittype, boundary_conditions, generator,
    codegen_option, lbm_algorithm, show_code).  `dtype='float32'` selects fp32 STORAGE of
    the populations (arithmetic stays fp64) -- new functionality, the reference ignores
    its dtype argument (simulation.py:89-91, storage.py:67).  `compute_dtype='float32'` (only with
    dtype='float32') also runs the time-step kernels in fp32 arithmetic: the all-single-precision
    mode, tolerance stated in tests/test_gpu_parity.py; moments (`sol.m`), initialisation and wall
    equilibria are always computed in fp64.
    `in_place=True` selects in-place streaming (AA pattern): ONE population array instead of two -- half
    the HBM footprint, same traffic per step, bit-identical results (tests/test_gpu_aa.py); single GPU.
    `slab=(rank, nranks)` + `nccl_id` run this process as one x-slab of a multi-GPU run (NCCL
    send/recv halo).  `gather(bytes) -> [bytes of every rank]` (an all-gather provided by the caller,
    e.g. torch.distributed.all_gather_object) additionally enables the direct NVLink halo: the fused
    kernel stores the slab-face populations straight into the neighbours' ghost planes.
    """

    def __init__(self, dico, sorder=None, dtype="float64", check_inverse=False, initialize=True,
                 slab=None, nccl_id=None, gather=None, compute_dtype=None, in_place=False):
        generator = str(dico.get("generator", "cuda")).upper()
        if generator != "CUDA":
            raise ValueError(
                "pylbm_b200 only provides generator='cuda' (got %r); there is no CPU fallback" % generator
            )
        rt.ensure_gpu()
        storage, compute = self._storage_names(dtype, compute_dtype)
        self._engine_defaults(storage, compute, slab, nccl_id, gather, in_place)

        rank, nranks = slab if slab is not None else (0, 1)
        topo = None
        if nranks > 1:
            from .stencil import Stencil

            topo = SlabTopology(Stencil.extract_dim(dico), rank, nranks)
            topo.gather = gather          # used by H5File to bring the slabs to rank 0
        self.domain = Domain(dico, need_validation=False, topology=topo)
        self.scheme = Scheme(dico, check_inverse=check_inverse, need_validation=False)
        if self.domain.dim != self.scheme.dim:
            raise ValueError("Solution: the dimension of the domain and of the scheme are not the same")

        self._update_m = True
        self.t = 0.0
        self.nt = 0
        self.dt_ = self.domain.dx / self.scheme.la
        self.dim = self.domain.dim
        self.extra_parameters = {}

        # ---- generated kernels -----------------------------------------
        user_algo = dico.get("lbm_algorithm", None) or {}
        codegen_opt = dico.get("codegen_option", None)
        want_source = bool(dico.get("show_code", False) or (codegen_opt and codegen_opt.get("directory")))
        # (PYLBM_B200_AA_LIBRARY=1: a two-array run on the library generated for in-place streaming, so that
        # both variants come from ONE lowering -- sympy.cse groups sums differently from call to call, which
        # moves last bits between separately generated libraries)
        self._aa_library = self.in_place or bool(os.environ.get("PYLBM_B200_AA_LIBRARY"))
        self.algo, lib_path, source = build_kernel_library(self.scheme, user_algo.get("settings", {}), storage,
                                                           need_source=want_source, compute=compute,
                                                           aa=self._aa_library)
        if dico.get("show_code", False):
            print(source)
        if codegen_opt and codegen_opt.get("directory"):
            outdir = os.path.realpath(codegen_opt["directory"])
            os.makedirs(outdir, exist_ok=True)
            with open(os.path.join(outdir, os.path.basename(lib_path)[3:-3] + ".cu"), "w") as fh:
                fh.write(source)
        self.kernels = rt.KernelLibrary(lib_path)
        self._algo_settings = user_algo.get("settings", {})
        self.generator = types.SimpleNamespace(backend="CUDA", module=self.kernels)

        # ---- storage ----------------------------------------------------
        self.container = CudaContainer(self.domain, self.scheme, sorder, storage, self.in_place)

        # ---- boundary lists ---------------------------------------------
        self.bc = Boundary(self.domain, self.generator, dico)
        for method in self.bc.methods:
            method.set_iload()

        self.init_type = dico.get("inittype", "moments")
        self.init_data = dico.get("init", None)
        self._need_init = True
        if initialize:
            self._initialize()

    def _build_kernels(self, storage, compute):
        return rt.KernelLibrary(build_kernel_library(self.scheme, self._algo_settings, storage, compute=compute,
                                                     aa=self._aa_library)[1])
