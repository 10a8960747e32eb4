"""
ORACLE -- test infrastructure, NOT part of the product.

CPU restatement of the reference's `one_time_step` path (pylbm v0.11.0, Cython
generator) used only as the checker in tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py.  Nothing under
pylbm_b200/ imports it.

Parity status: PINNED.  This restatement is checked (tests/test_oracle_golden.py)
against fixtures produced by running the unmodified reference in the build
container (tools/make_golden.py -> tests/golden/*.npz: boundary lists, moment
matrices and conserved moments after N steps with the reference's Cython
generator) and against the reference's own golden HDF5 fields
(tests/reference/*.h5, converted by the same script).

What is restated, with the reference lines it follows:

  OracleSimulation.one_time_step   pylbm/simulation.py:392-420
    periodic_update                pylbm/storage.py:333-367 (one rank: ghost <- opposite
                                   interior planes, dimension by dimension)
    boundary loops                 pylbm/boundary.py:462-464 (bounce_back), 608-618
                                   (Bouzidi_bounce_back on a copy `fcopy`, 536-559), 678-680
                                   (anti_bounce_back), 745-756 (Bouzidi_anti_bounce_back),
                                   818 (neumann*)  -- sequential, in list order
    fused pull kernel              pylbm/algorithm/pull.py:11-59 with base.py:298-334 (f2m,
                                   relative velocity), 404-428 (relaxation), 255-263 (restore
                                   conserved moments), 439-495 + ode.py:11-16 (source terms),
                                   346-369 (m2f); dense (T(u) M) and (M^-1 T(-u)) matrices as
                                   the reference builds them (base.py:102-112)
    layout                         array-of-structures [x, y, z, Q], C order, loops over
                                   [vmax, n - vmax) (container.py:75-91, base.py:170-189)
    arithmetic                     generated C, gcc -O2 -ffp-contract=off: IEEE double, no FMA,
                                   element-wise sequential in-place matrix assignments
                                   (generator/printing/cython.py:316-350)
  initialization / wall equilibrium  simulation.py:258-320, boundary.py:238-305

The problem description (stencil, M, equilibria, boundary lists) comes from the
host front-end of the package, which is itself pinned bit-for-bit against the
reference by the same fixtures.

Since round 2 the unmodified reference itself is a second, fully independent checker:
tools/make_ref.sh installs it into oracle/_ref/ (git-ignored, travels to the GPU box), where
tests/test_gpu_plugin.py and tests/test_gpu_standalone.py run its Cython and NumPy generators
live beside the CUDA path, and bench.py times it as the CPU baseline (kind "reference").
"""

import ctypes
import hashlib
import os
import subprocess
import types

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")

rel_ux, rel_uy, rel_uz = sp.symbols("rel_ux, rel_uy, rel_uz", real=True)


# ---------------------------------------------------------------------------
# literal symbolic algorithm (reference: algorithm/base.py, algorithm/pull.py)
# ---------------------------------------------------------------------------
class _Printer(C99CodePrinter):
    def _print_Pow(self, expr):
        if expr.exp.is_Integer and 0 < int(expr.exp) <= 8:
            b = self.parenthesize(expr.base, 1000)
            return "(" + "*".join([b] * int(expr.exp)) + ")"
        if expr.exp == -1:
            return "(1.0/%s)" % self.parenthesize(expr.base, 1000)
        return super()._print_Pow(expr)

    def _print_Rational(self, expr):
        return repr(float(sp.Float(expr, 30)))

    def _print_Integer(self, expr):
        return "%d.0" % int(expr) if int(expr) >= 0 else "(%d.0)" % int(expr)

    def _print_Float(self, expr):
        v = repr(float(expr))
        return v if float(expr) >= 0 else "(%s)" % v


_pr = _Printer()


def _rsub(expr, repl):
    for _ in range(len(repl) + 1):
        new = expr.subs(repl)
        if new == expr:
            return new
        expr = new
    return expr


class LiteralAlgorithm:
    """sequential statements of every routine, written the way the reference writes them."""

    def __init__(self, scheme):
        self.scheme = scheme
        self.dim = scheme.dim
        ns = self.ns = int(scheme.stencil.nv_ptr[-1])
        self.nc = len(scheme.consm)
        self.m = [sp.Symbol("m[%d]" % i) for i in range(ns)]
        self.vel = scheme.stencil.get_all_velocities()
        params = [(sp.Symbol(str(k)), v) for k, v in scheme.param.items()] + list(scheme.param.items())
        full = params + [(k, self.m[int(i)]) for k, i in scheme.consm.items()]
        self.M, self.invM = scheme.M, scheme.invM
        self.eq = [_rsub(sp.sympify(e), full) for e in scheme.EQ]
        self.s = [_rsub(sp.sympify(e), full) for e in scheme.s]
        self.rel = scheme.rel_vel is not None
        if self.rel:
            self.rel_sym = [rel_ux, rel_uy, rel_uz][: self.dim]
            self.rel_vel = [_rsub(sp.sympify(e), full) for e in scheme.rel_vel]
            # dense polynomial matrices, as in base.py:109-112
            self.Mu = (scheme.Tu * scheme.M).applyfunc(sp.expand)
            self.invMu = (scheme.invM * scheme.Tmu).applyfunc(sp.expand)
            self.eq_u = list((scheme.Tu * sp.Matrix(self.eq)).applyfunc(sp.expand))
        self.source = []
        for src in scheme._source_terms:
            if src:
                for k, v in src.items():
                    self.source.append((_rsub(sp.sympify(k), full), _rsub(sp.sympify(v), full)))

    def _rows(self, mat, vec, rows):
        return [(i, sum((mat[i, j] * vec[j] for j in range(self.ns)), sp.Integer(0))) for i in rows]

    def _source(self):
        dt = sp.Symbol("dt")
        return [(lhs, lhs + dt / 2 * rhs) for lhs, rhs in self.source]

    def fused(self, f):
        """list of (lhs, rhs) in order + list of fnew expressions."""
        m, nc, ns = self.m, self.nc, self.ns
        stm = []
        if self.rel:
            stm += [(m[i], e) for i, e in self._rows(self.M, f, range(nc))]
            stm += [(self.rel_sym[d], self.rel_vel[d]) for d in range(self.dim)]
            stm += [(m[i], e) for i, e in self._rows(self.Mu, f, range(nc, ns))]
        else:
            stm += [(m[i], e) for i, e in self._rows(self.M, f, range(ns))]
        stm += self._source()
        eq = self.eq_u if self.rel else self.eq
        for i in range(ns):
            if self.s[i] != 0:
                stm.append((m[i], (1 - self.s[i]) * m[i] + self.s[i] * eq[i]))
        if self.rel:
            stm += [(m[i], e) for i, e in self._rows(self.Mu, f, range(nc))]
        stm += self._source()
        inv = self.invMu if self.rel else self.invM
        out = [e for _, e in self._rows(inv, m, range(ns))]
        return stm, out

    def equilibrium(self):
        return [(self.m[i], self.eq[i]) for i in range(self.ns) if self.eq[i] != self.m[i]]


# ---------------------------------------------------------------------------
# C source
# ---------------------------------------------------------------------------
_C_TEMPLATE = r"""
#include <string.h>
#include <stdlib.h>
#define Q %(q)d
#define IDX(ix,iy,iz) ((((long)(ix))*ny + (iy))*nz + (iz))

/* fused pull stream + collide on the inner cells, AoS [x][y][z][Q] */
void one_time_step(const double* f, double* fnew, int nx, int ny, int nz, double t, double dt, const double* extra)
{
    const int vx0 = %(vmax0)d, vy0 = %(vmax1)d, vz0 = %(vmax2)d;
    #pragma omp parallel for schedule(static)
    for (int ix = vx0; ix < nx - vx0; ++ix)
    for (int iy = vy0; iy < ny - vy0; ++iy)
    for (int iz = vz0; iz < nz - vz0; ++iz) {
        double m[Q]; double rel_ux = 0, rel_uy = 0, rel_uz = 0;
        (void)rel_ux; (void)rel_uy; (void)rel_uz; (void)t; (void)dt; (void)extra;
%(fused)s
    }
}

void f2m(const double* f, double* mm, int nx, int ny, int nz)
{
    for (long c = 0; c < (long)nx*ny*nz; ++c) {
        const double* ff = f + c*Q; double* m = mm + c*Q;
%(f2m)s
    }
}

void m2f(const double* mm, double* f, int nx, int ny, int nz)
{
    for (long c = 0; c < (long)nx*ny*nz; ++c) {
        const double* m = mm + c*Q; double* ff = f + c*Q;
%(m2f)s
    }
}

void equilibrium(double* mm, int nx, int ny, int nz, double t, double dt, const double* extra)
{
    for (long c = 0; c < (long)nx*ny*nz; ++c) {
        double* m = mm + c*Q;
        (void)t; (void)dt; (void)extra;
%(equilibrium)s
    }
}

/* periodic ghost update, one rank: dimension by dimension; skip_axis (or -1) is left to the caller
   (slab decomposition: that axis is exchanged with the neighbour ranks first) */
void periodic_update(double* f, int nx, int ny, int nz, int skip_axis)
{
    const int w[3] = {%(vmax0)d, %(vmax1)d, %(vmax2)d};
    const int n[3] = {nx, ny, nz};
    for (int d = 0; d < 3; ++d) {
        if (w[d] == 0 || d == skip_axis) continue;
        for (int ix = 0; ix < nx; ++ix) for (int iy = 0; iy < ny; ++iy) for (int iz = 0; iz < nz; ++iz) {
            int i[3] = {ix, iy, iz};
            int src = -1;
            if (i[d] < w[d]) src = n[d] - 2*w[d] + i[d];
            else if (i[d] >= n[d] - w[d]) src = i[d] - (n[d] - w[d]) + w[d];
            if (src < 0) continue;
            int s[3] = {ix, iy, iz}; s[d] = src;
            memcpy(f + IDX(ix,iy,iz)*Q, f + IDX(s[0],s[1],s[2])*Q, Q*sizeof(double));
        }
    }
}

/* boundary loops: index rows are [k, ix, iy, iz] (missing dimensions = 0), sequential */
#define AT(a, r) a[IDX(r[1], r[2], r[3])*Q + r[0]]
void bc_bounce_back(double* f, const int* is, const int* l0, const double* rhs, long n, int nx, int ny, int nz)
{ for (long i = 0; i < n; ++i) { const int* s = is+4*i; const int* a = l0+4*i; AT(f,s) = AT(f,a) + rhs[i]; } }
void bc_anti_bounce_back(double* f, const int* is, const int* l0, const double* rhs, long n, int nx, int ny, int nz)
{ for (long i = 0; i < n; ++i) { const int* s = is+4*i; const int* a = l0+4*i; AT(f,s) = -AT(f,a) + rhs[i]; } }
void bc_bouzidi_bounce_back(double* f, const double* fcopy, const int* is, const int* l0, const int* l1,
                            const double* rhs, const double* dist, long n, int nx, int ny, int nz)
{ for (long i = 0; i < n; ++i) { const int* s = is+4*i; const int* a = l0+4*i; const int* b = l1+4*i;
    AT(f,s) = (1 - dist[i])*AT(fcopy,b) + dist[i]*AT(fcopy,a) + rhs[i]; } }
void bc_bouzidi_anti_bounce_back(double* f, const int* is, const int* l0, const int* l1,
                                 const double* rhs, const double* dist, long n, int nx, int ny, int nz)
{ for (long i = 0; i < n; ++i) { const int* s = is+4*i; const int* a = l0+4*i; const int* b = l1+4*i;
    AT(f,s) = (1 - dist[i])*AT(f,b) - dist[i]*AT(f,a) + rhs[i]; } }
void bc_neumann(double* f, const int* is, const int* l0, long n, int nx, int ny, int nz)
{ for (long i = 0; i < n; ++i) { const int* s = is+4*i; const int* a = l0+4*i; AT(f,s) = AT(f,a); } }
"""


def _source_for(scheme, extra_names):
    algo = LiteralAlgorithm(scheme)
    ns, dim = algo.ns, algo.dim
    vmax = list(scheme.stencil.vmax) + [0] * (3 - dim)
    fsym = [sp.Symbol("F%d" % k) for k in range(ns)]
    extra_map = {sp.Symbol(n): sp.Symbol("extra[%d]" % i) for i, n in enumerate(extra_names)}

    def pr(e):
        return _pr.doprint(sp.sympify(e).xreplace(extra_map))

    lines = []
    for k in range(ns):
        off = [-int(c) for c in algo.vel[k]] + [0] * (3 - dim)
        lines.append(
            "        const double F%d = f[IDX(ix + (%d), iy + (%d), iz + (%d))*Q + %d];" % (k, off[0], off[1], off[2], k)
        )
    stm, out = algo.fused(fsym)
    for lhs, rhs in stm:
        lines.append("        %s = %s;" % (lhs, pr(rhs)))
    for k, e in enumerate(out):
        lines.append("        fnew[IDX(ix, iy, iz)*Q + %d] = %s;" % (k, pr(e)))
    fused = "\n".join(lines)

    ff = [sp.Symbol("ff[%d]" % k) for k in range(ns)]
    f2m = "\n".join("        m[%d] = %s;" % (i, pr(e)) for i, e in algo._rows(algo.M, ff, range(ns)))
    m2f = "\n".join("        ff[%d] = %s;" % (i, pr(e)) for i, e in algo._rows(algo.invM, algo.m, range(ns)))
    equi = "\n".join("        %s = %s;" % (lhs, pr(rhs)) for lhs, rhs in algo.equilibrium())
    return _C_TEMPLATE % dict(
        q=ns, vmax0=vmax[0], vmax1=vmax[1], vmax2=vmax[2], fused=fused, f2m=f2m, m2f=m2f, equilibrium=equi
    )


def _free_extras(scheme):
    """user symbols that stay free in the kernels (given through extra_parameters)."""
    algo = LiteralAlgorithm(scheme)
    stm, out = algo.fused([sp.Symbol("F%d" % k) for k in range(algo.ns)])
    known = set(algo.m) | set(sp.Symbol("F%d" % k) for k in range(algo.ns)) | {sp.Symbol("dt"), sp.Symbol("t")}
    known |= {rel_ux, rel_uy, rel_uz}
    free = set()
    for _, rhs in stm:
        free |= sp.sympify(rhs).free_symbols
    for e in out:
        free |= sp.sympify(e).free_symbols
    return sorted(str(s) for s in free - known)


def build_library(scheme, openmp=False):
    os.makedirs(BUILD_DIR, exist_ok=True)
    extras = _free_extras(scheme)
    src = _source_for(scheme, extras)
    tag = hashlib.sha256((src + str(openmp)).encode()).hexdigest()[:16]
    lib = os.path.join(BUILD_DIR, "liboracle_%s.so" % tag)
    if not os.path.exists(lib):
        cfile = os.path.join(BUILD_DIR, "oracle_%s.c" % tag)
        with open(cfile, "w") as fh:
            fh.write(src)
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", lib + ".tmp", cfile, "-lm"]
        if openmp:
            cmd.insert(1, "-fopenmp")
        subprocess.run(cmd, check=True)
        os.replace(lib + ".tmp", lib)
    return ctypes.CDLL(lib), extras


# ---------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------
def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _AoS:
    """host array [x, y, z, Q] with the reference's [nv, x, y, z] view and moment keys."""

    def __init__(self, nv, nspace, consm):
        self.nspace = tuple(nspace)
        self.dim = len(nspace)
        full = tuple(nspace) + (1,) * (3 - self.dim)
        self.array = np.zeros(full + (nv,))
        self.consm = dict(consm)

    @property
    def swaparray(self):
        view = np.moveaxis(self.array, -1, 0)
        return view[(slice(None),) * (1 + self.dim) + (0,) * (3 - self.dim)]

    def _key(self, key):
        return self.consm[key] if isinstance(key, (sp.Symbol, sp.IndexedBase)) else key

    def __getitem__(self, key):
        return self.swaparray[self._key(key)]

    def __setitem__(self, key, values):
        self.swaparray[self._key(key)] = values


class OracleSimulation:
    """
    CPU twin of pylbm_b200.Simulation built on the same dictionary; exposes what the tests
    compare: `m[sym]`, `F`, boundary lists `bc.methods`, `one_time_step()`.
    """

    def __init__(self, dico, openmp=False, threads=None, topology=None, exchange=None):
        """`topology` (pylbm_b200.domain.SlabTopology) + `exchange(array, width)` run this object as
        one x-slab: `exchange` must fill the x ghost planes of the AoS array from the neighbour
        ranks (reference: storage.py:333-367 with several ranks); used by the gloo tests."""
        from pylbm_b200.boundary import Boundary
        from pylbm_b200.domain import Domain
        from pylbm_b200.scheme import Scheme

        self.domain = Domain(dico, topology=topology)
        self.exchange = exchange
        self.scheme = Scheme(dico)
        self.dim = self.domain.dim
        self.lib, self.extra_names = build_library(self.scheme, openmp=openmp)
        if openmp and threads:
            os.environ["OMP_NUM_THREADS"] = str(threads)
        self.nv = int(self.scheme.stencil.nv_ptr[-1])
        shape = self.domain.shape_halo
        self.n3 = list(shape) + [1] * (3 - self.dim)
        consm = self.scheme.consm
        self._m = _AoS(self.nv, shape, consm)
        self._F = _AoS(self.nv, shape, consm)
        self._Fnew = _AoS(self.nv, shape, consm)
        self.container = types.SimpleNamespace(nv=self.nv, m=self._m, F=self._F, Fnew=self._Fnew)
        self.t, self.nt = 0.0, 0
        la = self.scheme.la
        if isinstance(la, sp.Expr):
            la = float(la.subs(list(self.scheme.param.items())))
        self.dt = self.domain.dx / la
        self.extra_parameters = {}
        self._update_m = True

        self.bc = Boundary(self.domain, None, dico)
        for method in self.bc.methods:
            method.set_iload()
        self._initialize(dico)

    # ---- generated-kernel wrappers ---------------------------------------
    def _extra(self):
        vals = {str(k): float(v) for k, v in self.extra_parameters.items()}
        arr = np.array([vals[n] for n in self.extra_names] + [0.0])
        return arr

    def equilibrium(self, m_user=None):
        arr = self._m.array if m_user is None else np.ascontiguousarray(np.moveaxis(m_user.array, 0, -1))
        ncell = arr.size // self.nv
        ex = self._extra()
        self.lib.equilibrium(_ptr(arr), ctypes.c_int(ncell), ctypes.c_int(1), ctypes.c_int(1),
                             ctypes.c_double(self.t), ctypes.c_double(self.dt), _ptr(ex))
        if m_user is not None:
            m_user.array[...] = np.moveaxis(arr, -1, 0)

    def m2f(self, m_user=None, f_user=None):
        if m_user is None:
            src, dst = self._m.array, self._F.array
        else:
            src = np.ascontiguousarray(np.moveaxis(m_user.array, 0, -1))
            dst = np.empty_like(src)
        ncell = src.size // self.nv
        self.lib.m2f(_ptr(src), _ptr(dst), ctypes.c_int(ncell), ctypes.c_int(1), ctypes.c_int(1))
        if m_user is not None:
            f_user.array[...] = np.moveaxis(dst, -1, 0)

    def f2m(self):
        ncell = self._F.array.size // self.nv
        self.lib.f2m(_ptr(self._F.array), _ptr(self._m.array), ctypes.c_int(ncell), ctypes.c_int(1), ctypes.c_int(1))

    # ---- initialisation (reference: simulation.py:155-163, 258-320) ---------
    def _initialize(self, dico):
        coords = np.meshgrid(*self.domain.coords_halo, sparse=True, indexing="ij")
        init_type = dico.get("inittype", "moments")
        target = self._m if init_type == "moments" else self._F
        for k, v in (dico.get("init", None) or {}).items():
            if isinstance(v, tuple):
                extraargs = v[1] if len(v) == 2 else ()
                target[k] = v[0](*(tuple(coords) + tuple(extraargs)))
            elif isinstance(v, types.FunctionType):
                target[k] = v(*coords)
            else:
                target[k] = v
        if init_type == "moments":
            self.equilibrium()
            self.m2f()
        else:
            self.f2m()
        self._Fnew.array[...] = self._F.array
        for method in self.bc.methods:
            method.prepare_rhs(self)
            method.fix_iload()
            method.set_rhs()

    # ---- one time step (reference: simulation.py:373-420) -------------------
    def _rows4(self, idx):
        out = np.zeros((idx.shape[0], 4), dtype=np.int32)
        out[:, : idx.shape[1]] = idx
        return np.ascontiguousarray(out)

    def boundary_condition(self):
        f = self._F.array
        n = [ctypes.c_int(v) for v in self.n3]
        if self.exchange is not None:
            self.exchange(f, int(self.domain.stencil.vmax[0]))
            self.lib.periodic_update(_ptr(f), *n, ctypes.c_int(0))
        else:
            self.lib.periodic_update(_ptr(f), *n, ctypes.c_int(-1))
        for method in self.bc.methods:
            if method.is_time_dependent:
                method.update_feq(self)
            method.set_rhs()
            ist = self._rows4(method.istore)
            l0 = self._rows4(method.iload[0])
            cnt = ctypes.c_long(ist.shape[0])
            name = type(method).__name__
            if name == "BounceBack":
                self.lib.bc_bounce_back(_ptr(f), _ptr(ist), _ptr(l0), _ptr(method.rhs), cnt, *n)
            elif name == "AntiBounceBack":
                self.lib.bc_anti_bounce_back(_ptr(f), _ptr(ist), _ptr(l0), _ptr(method.rhs), cnt, *n)
            elif name == "BouzidiBounceBack":
                fcopy = f.copy()
                l1 = self._rows4(method.iload[1])
                self.lib.bc_bouzidi_bounce_back(_ptr(f), _ptr(fcopy), _ptr(ist), _ptr(l0), _ptr(l1),
                                                _ptr(method.rhs), _ptr(method.s), cnt, *n)
            elif name == "BouzidiAntiBounceBack":
                l1 = self._rows4(method.iload[1])
                self.lib.bc_bouzidi_anti_bounce_back(_ptr(f), _ptr(ist), _ptr(l0), _ptr(l1),
                                                     _ptr(method.rhs), _ptr(method.s), cnt, *n)
            elif name.startswith("Neumann"):
                self.lib.bc_neumann(_ptr(f), _ptr(ist), _ptr(l0), cnt, *n)
            else:
                raise NotImplementedError(name)

    def one_time_step(self):
        self._update_m = True
        self.boundary_condition()
        ex = self._extra()
        n = [ctypes.c_int(v) for v in self.n3]
        self.lib.one_time_step(_ptr(self._F.array), _ptr(self._Fnew.array), *n,
                               ctypes.c_double(self.t), ctypes.c_double(self.dt), _ptr(ex))
        self._F, self._Fnew = self._Fnew, self._F
        self.container.F, self.container.Fnew = self._F, self._Fnew
        self.t += self.dt
        self.nt += 1

    # ---- results --------------------------------------------------------------
    def _inner(self):
        return tuple(slice(v, -v) if v > 0 else slice(None) for v in self.domain.stencil.vmax)

    @property
    def m(self):
        sim = self

        class _Get:
            def __getitem__(self_, key):
                if sim._update_m:
                    sim._update_m = False
                    sim.f2m()
                return sim._m[key][sim._inner()]

        return _Get()

    @property
    def m_halo(self):
        sim = self

        class _Get:
            def __getitem__(self_, key):
                if sim._update_m:
                    sim._update_m = False
                    sim.f2m()
                return sim._m[key]

        return _Get()

    @property
    def F_halo(self):
        return self._F
