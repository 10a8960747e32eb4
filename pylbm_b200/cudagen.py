"""
Lowering of the per-cell kernel IR (algorithm.py) to CUDA C for sm_100a.

This is the `generator='cuda'` code generator: it plays the role the
reference's CythonCodeGen + CythonCodePrinter play for `generator='cython'`
(reference: pylbm/generator/codegen.py:1014-1178, printing/cython.py:72-435),
but emits hand-written-template CUDA C instead of Cython loops:

* one thread per lattice cell, 128-thread blocks laid along the fastest axis so
  that every population load/store of a warp is one coalesced 256-byte access;
* device arrays are SoA [population][x][y][z] with padded rows whose first
  interior cell is 128-byte aligned (see storage.py / include/lbmk.h);
* the Q×Q moment transforms, equilibria and relaxation are straight-line
  register code after SSA renaming + common-subexpression elimination, with
  exact rational coefficients printed as round-trip double literals; nvcc
  contracts a*b+c into DFMA.  No tensor cores: the step is HBM-bound
  (2·Q·8 bytes per cell for a few hundred flops).

The generated translation unit exports, per routine, a C-ABI launcher
`int lbmk_<routine>(const void* in, void* out, const lbmk_grid* g,
const double* scalars, void* stream)` plus `lbmk_describe` (include/lbmk.h).
"""

import hashlib

import sympy as sp
from sympy.printing.c import C99CodePrinter

__all__ = ["generate_source", "CudaPrinter", "lower_statements"]

ABI_VERSION = 1


class CudaPrinter(C99CodePrinter):
    """C printer for the per-cell arithmetic: integer powers as products, reciprocal as 1.0/x,
    rationals and floats as round-trip double literals (`single=True`: float literals, the decimal
    text of the double followed by `f`, and the float math functions)."""

    def __init__(self, single=False):
        super().__init__()
        self._sfx = "f" if single else ""

    def _literal(self, value):
        text = repr(float(value)) + self._sfx
        return text if float(value) >= 0 else "(%s)" % text

    def _print_Rational(self, expr):
        return self._literal(float(sp.Float(expr, 30)))

    def _print_Integer(self, expr):
        return self._literal(int(expr))

    def _print_Float(self, expr):
        return self._literal(float(expr))

    def _print_Pow(self, expr):
        base, exp = expr.base, expr.exp
        one = "1.0" + self._sfx
        if exp.is_Integer:
            n = int(exp)
            b = self.parenthesize(base, 1000)
            if n == -1:
                return "(%s/%s)" % (one, b)
            if 0 < n <= 8:
                return "(" + "*".join([b] * n) + ")"
            if -8 <= n < 0:
                return "(%s/(" % one + "*".join([b] * (-n)) + "))"
        if exp == sp.Rational(1, 2):
            return "sqrt%s(%s)" % (self._sfx, self._print(base))
        if exp == sp.Rational(-1, 2):
            return "rsqrt%s(%s)" % (self._sfx, self._print(base))
        return "pow%s(%s, %s)" % (self._sfx, self._print(base), self._print(exp))


class ExplicitPrinter(CudaPrinter):
    """Sums and products as explicit round-to-nearest intrinsics: `fma(a, b, acc)` for every product term
    of a sum, `__dadd_rn` / `__dmul_rn` otherwise.  nvcc can then neither contract nor leave uncontracted
    anything on its own, so every instantiation of a kernel (two-array, walls, in-place even / odd step)
    performs bit for bit the same arithmetic -- with the infix form the even-step instantiation of the
    D3Q19 kernel came out with 4 of its 19 outputs contracted differently (last-bit differences).  The
    dependency chains are the ones of the infix form (left to right).  Used by the libraries generated
    for in-place streaming."""

    def __init__(self, single=False):
        super().__init__(single)
        self._fma = "__fmaf_rn" if single else "__fma_rn"
        self._mul = "__fmul_rn" if single else "__dmul_rn"
        self._add = "__fadd_rn" if single else "__dadd_rn"

    def _product(self, factors):
        text = self._print(factors[0])
        for f in factors[1:]:
            text = "%s(%s, %s)" % (self._mul, text, self._print(f))
        return text

    def _print_Mul(self, expr):
        coeff, factors = expr.as_coeff_mul()
        factors = list(factors)
        if not factors:
            return self._print(coeff)
        if coeff == 1:
            return self._product(factors)
        if coeff == -1:
            return "(-%s)" % self._product(factors)
        return "%s(%s, %s)" % (self._mul, self._literal(float(sp.Float(coeff, 30))), self._product(factors))

    def _print_Add(self, expr):
        terms = self._as_ordered_terms(expr, order=None)
        acc = self._print(terms[0])
        for term in terms[1:]:
            coeff, factors = term.as_coeff_mul()
            factors = list(factors)
            if not factors:                                   # a number
                acc = "%s(%s, %s)" % (self._add, acc, self._print(term))
            elif len(factors) == 1 and coeff == 1:
                acc = "%s(%s, %s)" % (self._add, acc, self._print(factors[0]))
            elif len(factors) == 1 and coeff == -1:
                acc = "%s(%s, -%s)" % (self._add, acc, self._print(factors[0]))
            elif len(factors) == 1:
                acc = "%s(%s, %s, %s)" % (self._fma, self._literal(float(sp.Float(coeff, 30))),
                                          self._print(factors[0]), acc)
            else:
                head = self._product(factors[:-1])
                if coeff == -1:
                    head = "(-%s)" % head
                elif coeff != 1:
                    head = "%s(%s, %s)" % (self._mul, self._literal(float(sp.Float(coeff, 30))), head)
                acc = "%s(%s, %s, %s)" % (self._fma, head, self._print(factors[-1]), acc)
        return acc


_printers = {False: CudaPrinter(False), True: CudaPrinter(True)}
_explicit_printers = {False: ExplicitPrinter(False), True: ExplicitPrinter(True)}


def _to_exact(expr):
    """Floats -> rationals so that CSE / expansion work on exact coefficients."""
    expr = sp.sympify(expr)
    floats = expr.atoms(sp.Float)
    if not floats:
        return expr
    return expr.xreplace({f: sp.Rational(float(f)) for f in floats})


def lower_statements(statements, outputs, cse=True):
    """
    Sequential in-place statements -> SSA -> CSE.
    Returns (list of (name, expr) temporaries in evaluation order, list of
    output expressions).
    """
    current = {}
    ssa = []
    counter = {}
    for lhs, rhs in statements:
        rhs = _to_exact(sp.sympify(rhs).xreplace(current))
        n = counter.get(lhs, 0)
        counter[lhs] = n + 1
        new = sp.Symbol("%s_%d" % (lhs, n), real=True)
        current[lhs] = new
        ssa.append((new, rhs))
    outs = [_to_exact(sp.sympify(o).xreplace(current)) for o in outputs]

    if not cse:
        return ssa, outs

    # one reciprocal per distinct base: b**-n -> (1/b)**n, so that e.g. qx/rho, qy/rho and
    # qx**2/rho**2 share a single fp64 division
    recips = {}

    def _recip(expr):
        def repl(p):
            base, n = p.base, -int(p.exp)
            if base not in recips:
                recips[base] = sp.Symbol("inv_%d" % len(recips), real=True)
            return recips[base] ** n
        return expr.replace(lambda e: e.is_Pow and e.exp.is_Integer and e.exp < 0, repl)

    ssa = [(lhs, _recip(rhs)) for lhs, rhs in ssa]
    outs = [_recip(o) for o in outs]
    ssa = ssa + [(sym, 1 / base) for base, sym in recips.items()]

    # forward-substitute trivial copies / keep the statement structure for CSE
    exprs = [rhs for _, rhs in ssa] + outs
    names = sp.numbered_symbols("t_", real=True)
    repl, reduced = sp.cse(exprs, symbols=names, optimizations="basic", order="none")
    temps = list(repl)
    for (lhs, _), red in zip(ssa, reduced[: len(ssa)]):
        temps.append((lhs, red))
    # CSE temporaries may reference SSA names defined "later" in `temps`
    # (cse pulls sub-expressions to the front); order them topologically.
    temps = _topological(temps)
    return temps, reduced[len(ssa) :]


def _topological(assignments):
    defined = {lhs: rhs for lhs, rhs in assignments}
    order, state = [], {}

    def visit(sym):
        st = state.get(sym, 0)
        if st == 2:
            return
        if st == 1:
            raise ValueError("cyclic dependency in kernel statements at %s" % sym)
        state[sym] = 1
        for dep in defined[sym].free_symbols:
            if dep in defined:
                visit(dep)
        state[sym] = 2
        order.append((sym, defined[sym]))

    import sys

    limit = sys.getrecursionlimit()
    sys.setrecursionlimit(max(limit, 20000))
    try:
        for lhs, _ in assignments:
            visit(lhs)
    finally:
        sys.setrecursionlimit(limit)
    return order


def count_ops(temps, outs):
    """(adds, muls, divs) of the lowered kernel, for the DESIGN roofline notes."""
    add = mul = div = 0
    for e in [r for _, r in temps] + list(outs):
        for node in sp.preorder_traversal(e):
            if node.is_Add:
                add += len(node.args) - 1
            elif node.is_Mul:
                mul += len(node.args) - 1
            elif node.is_Pow:
                if node.exp.is_Integer and node.exp > 0:
                    mul += int(node.exp) - 1
                else:
                    div += 1
    return add, mul, div


_HEADER = r"""// Generated by pylbm_b200.cudagen -- do not edit.  sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define LBMK_ABI_VERSION %(abi)d
#define LBMK_WRAP_PDL 0x100
// programmatic dependent launch: the next kernel of the stream may be scheduled once every block of this
// one has started; nothing is read or written before the previous kernel has completed (no-ops without
// the launch attribute)
#define LBMK_PDL_PROLOGUE()                                        \
    asm volatile("griddepcontrol.launch_dependents;");             \
    asm volatile("griddepcontrol.wait;" ::: "memory")

extern "C" {
typedef struct {
    int n[3];          // halo-inclusive logical sizes, canonical 3-D (slowest .. fastest)
    int lo[3];         // loop begin per axis
    int hi[3];         // loop end per axis
    int tx;            // threads of a block along the fastest axis (power of two <= block size)
    int w[3];          // ghost width per axis
    int wrap;          // bit a set: the fused kernel also writes the periodic images along axis a
    int64_t pitch;     // elements between two consecutive rows of the fastest axis
    int64_t lead;      // position of logical index 0 inside a row
    int64_t pstride;   // elements between two populations
} lbmk_grid;
typedef struct {
    void* lo;            // neighbour array receiving the images of my LOW-face cells (its high ghost)
    void* hi;            // neighbour array receiving the images of my HIGH-face cells (its low ghost)
    int64_t pstride_lo;  // population stride of those arrays
    int64_t pstride_hi;
    int nin_lo;          // interior size of the `lo` neighbour along the slab axis
} lbmk_peers;
}
typedef struct {       // per-axis description of the cells that own a periodic / neighbour image
    int below[3];      // index <  below: image at +dlow  (in a high ghost layer)
    int above[3];      // index >= above: image at dhigh (< 0, in a low ghost layer)
    long long dlow[3];
    long long dhigh[3];
} lbmk_images;
typedef struct {       // bounce-back walls normal to the fastest axis handled by the fused kernel itself
    int lo_plane;      // index (fastest axis) of the cells next to the LOW wall
    int hi_plane;      // ... next to the HIGH wall
    int neg_lo;        // 0: bounce-back (+f), 1: anti bounce-back (-f)
    int neg_hi;
    double rhs[64];    // right-hand side per LOADED population (the one moving towards the wall)
} lbmk_walls;
extern "C" {
typedef struct {       // boundary entries evaluated by the pulling thread (see lbmk.h)
    const int* block_ptr;           // [nblocks + 1] first task of every 128-thread block
    const unsigned* code;           // thread | population << 8 | kind << 16
    const long long* l0;            // element positions read in the INPUT array
    const long long* l1;
    const double* const* rhs;       // address of the right-hand side of the entry (NULL: none)
    const double* dist;             // Bouzidi coefficient
    int ngroups_y, ngroups_x, tx;   // launch geometry the table was built for
} lbmk_tasks;
}
#define SLAB %(slab)d

typedef %(storage)s real_f;   // storage type of the populations in HBM
// a library generated for in-place streaming launches the fused kernel with fin == fout: no __restrict__,
// no non-coherent loads there
#if %(aa)d
#define LBMK_RESTRICT
#define LBMK_LDG(p) __ldcg(p)     /* ld.global.cg: a global (not generic) load that is coherent at L2 */
#else
#define LBMK_RESTRICT __restrict__
#define LBMK_LDG(p) __ldg(p)
#endif
// fp32 -> fp64 on the integer pipe (exact for normal numbers: rebias the exponent, widen the mantissa);
// zero, subnormals, inf and nan take the conversion instruction.  Experiment: PYLBM_B200_F2D=int.
__device__ __forceinline__ double lbmk_f2d(float x) {
    const unsigned u = __float_as_uint(x);
    const unsigned e = (u >> 23) & 0xffu;
    const unsigned hi = (u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u);
    double d = __hiloint2double((int)hi, (int)(u << 29));
    if (e == 0u || e == 255u) d = (double)x;
    return d;
}
__device__ __forceinline__ double lbmk_f2d(double x) { return x; }
typedef double real_m;        // moments are always stored in fp64
#define LBMK_BLOCK 128
"""

_KERNEL = r"""
// ---------------------------------------------------------------------------
// %(name)s : %(nin)d loads, %(nout)d stores per cell; %(ops)s
// ---------------------------------------------------------------------------
// byte offsets of the loads / stores relative to the cell, one per population: computed on the
// host by the launcher and read from the kernel-parameter constant bank, so that an access costs
// one 64-bit add (2 instructions) instead of re-deriving k*pstride + neighbour offset per thread
struct lbmk_offs_%(name)s {
    long long in[%(nin)d];
    long long out[%(nout)d];
    long long wall[%(nout)d];   // fused kernel with walls: store position of the bounced population
    long long nat[%(nout)d];    // population k of the cell itself (in-place even step with walls)
    unsigned fold;   // 0: blockIdx.y / blockIdx.z are the row group / the index of axis 0; otherwise the
                     // number of row groups per axis-0 index, (z, y) being one linear index (more than
                     // 65535 row groups or planes: e.g. a 2-D lattice with 100 000 rows)
    unsigned pair;   // a thread that computes several cells takes them along axis 0 (0) or from consecutive
                     // row groups of axis 1 (1: lattices with a single plane)
    unsigned vgy;    // row groups per plane of the lattice (the grid has fewer when pair == 1)
};

%(template)s__global__ void __launch_bounds__(LBMK_BLOCK, %(minblocks)d)
lbmk_kernel_%(name)s(const %(tin)s* %(restrict)s fin, %(tout)s* %(restrict)s fout, const lbmk_grid g,
    const lbmk_offs_%(name)s offs%(peer_param)s%(scalar_params)s)
{
    LBMK_PDL_PROLOGUE();
%(mode_decl)s    constexpr int NQ_ = %(nin)d; (void)NQ_;
    typedef %(tc)s real_c;   // arithmetic type of this kernel
    // 3-D grid: x = chunk of the fastest axis, y = group of `ty` rows of axis 1, z = index of axis 0
    // (no integer division in the prologue; tx is a power of two)
    const unsigned tid = threadIdx.x;
    const unsigned tx = (unsigned)g.tx;
    const unsigned txshift = 31u - (unsigned)__clz(tx);
    const unsigned ty = LBMK_BLOCK >> txshift;
    const int i2 = g.lo[2] + (int)(blockIdx.x * tx + (tid & (tx - 1)));
    unsigned by = blockIdx.y, bz = blockIdx.z;
    if (offs.fold) {                       // uniform branch, not taken for ordinary shapes
        const unsigned long long lin = (unsigned long long)blockIdx.z * gridDim.y + blockIdx.y;
        bz = (unsigned)(lin / offs.fold);
        by = (unsigned)(lin - (unsigned long long)bz * offs.fold);
    }
    const long long rowstride = g.pitch;
    const long long planestride = (long long)g.n[1] * g.pitch;
    (void)rowstride; (void)planestride;
%(thread_body)s
}

%(launch_head)s
{
    const int n0 = g->hi[0] - g->lo[0], n1 = g->hi[1] - g->lo[1], n2 = g->hi[2] - g->lo[2];
    if (n0 <= 0 || n1 <= 0 || n2 <= 0) return 0;
    if (g->tx <= 0 || (g->tx & (g->tx - 1)) || g->tx > LBMK_BLOCK) return -3;
    const int ty = LBMK_BLOCK / g->tx;
    dim3 grid((unsigned)((n2 + g->tx - 1) / g->tx), (unsigned)((n1 + ty - 1) / ty), (unsigned)n0);
    static const int noff[%(nin)d][3] = {%(offset_table)s};
    lbmk_offs_%(name)s offs;
    offs.fold = 0;
    offs.vgy = grid.y;
    offs.pair = (n0 > 1) ? 0u : 1u;
    if (offs.pair == 0u) grid.z = (grid.z + %(cpt)d - 1) / %(cpt)d;
    else grid.y = (grid.y + %(cpt)d - 1) / %(cpt)d;
    if (grid.y > 65535u || grid.z > 65535u) {
        // (row group, plane) as one linear index spread over (y, z)
        const unsigned long long total = (unsigned long long)grid.y * grid.z;
        offs.fold = grid.y;
        grid.y = 32768u;
        const unsigned long long gz = (total + grid.y - 1) / grid.y;
        if (gz > 65535ull) return -2;
        grid.z = (unsigned)gz;
    }
    const long long plane = (long long)g->n[1] * g->pitch;
    static const int inpop_[%(nin)d] = {%(inpop_table)s};   // population read as input k (identity, or the
                                                          // opposite population for a swapped in-place array)
    for (int k = 0; k < %(nin)d; ++k)
        offs.in[k] = (inpop_[k] * g->pstride + noff[k][0] * plane + noff[k][1] * g->pitch + noff[k][2]) * (long long)sizeof(%(tin)s);
    for (int k = 0; k < %(nout)d; ++k)
        offs.out[k] = k * g->pstride * (long long)sizeof(%(tout)s);
    for (int k = 0; k < %(nout)d; ++k) { offs.wall[k] = 0; offs.nat[k] = offs.out[k]; }
%(images_launch)s%(kernel_call)s
    return -(int)cudaGetLastError();
}
%(launch_tail)s"""


# Periodic images.  The reference refreshes the ghost layers at the START of every step by copying
# the opposite interior planes, axis by axis (storage.py:333-367).  Here the ghost values that can
# ever be read are produced at the END of the previous step instead: a cell within `w` of a face also
# stores its new populations to its periodic image(s), so no separate copy kernels run.
# A ghost cell displaced from the interior along the axes A is only ever read -- by the pull of the
# fused kernel or by a boundary-kernel load (bounce-back/Bouzidi/Neumann read `store + v` style
# positions) -- for the populations whose velocity points from that ghost cell into the interior along
# every axis of A.  Only those (population, image) pairs are stored; the other ghost entries are
# scratch (they are also never refreshed in the reference's Fnew, simulation.py:417).
_IMAGES_PROLOGUE = r"""    // ---- periodic / neighbour images of this cell (see the comment above _IMAGES_PROLOGUE) ----
    // offset of the image along each axis (0: none); thresholds and offsets come precomputed from
    // the launcher (an inactive axis has thresholds that never match)
    const long long d0 = i0 < img.below[0] ? img.dlow[0] : (i0 >= img.above[0] ? img.dhigh[0] : 0LL);
    const long long d1 = i1 < img.below[1] ? img.dlow[1] : (i1 >= img.above[1] ? img.dhigh[1] : 0LL);
    // (with walls on both faces of the fastest axis there is no periodic image along it)
    const long long d2 = WALLZ ? 0LL : (i2 < img.below[2] ? img.dlow[2] : (i2 >= img.above[2] ? img.dhigh[2] : 0LL));
    const bool wlo = WALLZ && i2 == walls.lo_plane, whi = WALLZ && i2 == walls.hi_plane;
    (void)wlo; (void)whi;
%(aa_shifts)s
    // array receiving the images that cross the SLAB axis (a neighbour's array with the peer halo)
    %(tout)s* qbase = fout;
    long long qps = g.pstride;
    if (pr.lo != nullptr) {
        const long long ds_ = (SLAB == 0) ? d0 : ((SLAB == 1) ? d1 : d2);
        if (ds_ > 0) { qbase = (%(tout)s*)pr.lo; qps = pr.pstride_lo; }
        else if (ds_ < 0) { qbase = (%(tout)s*)pr.hi; qps = pr.pstride_hi; }
    }
    // an image in a HIGH ghost layer (d > 0) is read by populations moving in -axis, and vice versa
    const bool p0 = d0 < 0, m0 = d0 > 0, p1 = d1 < 0, m1 = d1 > 0, p2 = d2 < 0, m2 = d2 > 0;
    (void)p0; (void)m0; (void)p1; (void)m1; (void)p2; (void)m2; (void)qbase; (void)qps;
"""

# launcher side of the above: per axis, cells with index < below have an image in the HIGH ghost
# layer at +dlow, cells with index >= above have one in the LOW ghost layer at dhigh (< 0)
_IMAGES_LAUNCH = r"""
    lbmk_images img;
    {
        const lbmk_peers* pp = peers;
        const long long stride_[3] = {(long long)g->n[1] * g->pitch, g->pitch, 1};
        for (int a = 0; a < 3; ++a) {
            const bool peer_ = (a == SLAB) && pp && pp->lo != nullptr;
            const bool on = ((((g->wrap >> a) & 1) != 0) || peer_) && g->w[a] > 0;
            const int nin = g->n[a] - 2 * g->w[a];
            img.below[a] = on ? 2 * g->w[a] : -2147483647 - 1;
            img.above[a] = on ? nin : 2147483647;
            img.dlow[a] = on ? (long long)(peer_ ? pp->nin_lo : nin) * stride_[a] : 0;
            img.dhigh[a] = on ? -(long long)nin * stride_[a] : 0;
        }
        // bounced population sym(k) at the neighbour c + v_k = c - noff[k] (walls; in-place even step)
        static const int sym_[%(nout)d] = {%(sym_table)s};
        for (int k = 0; k < %(nout)d; ++k)
            offs.wall[k] = (sym_[k] * g->pstride - (noff[k][0] * stride_[0] + noff[k][1] * stride_[1] + noff[k][2]))
                           * (long long)sizeof(%(tout)s);
        // in-place streaming (AA pattern): the even step stores population k into the slot it read for the
        // opposite population, (sym k, c + v_k); the odd step reads population k from (sym k, c)
        if (aa_phase == 1) for (int k = 0; k < %(nout)d; ++k) offs.out[k] = offs.wall[k];
        if (aa_phase == 2) for (int k = 0; k < %(nout)d; ++k) offs.in[k] = sym_[k] * g->pstride * (long long)sizeof(%(tout)s);
    }
"""


_THREAD_PLAIN = r"""    const int i1 = g.lo[1] + (int)(by * ty + (tid >> txshift));
    const int i0 = g.lo[0] + (int)bz;
    if (i2 >= g.hi[2] || i1 >= g.hi[1] || i0 >= g.hi[0]) return;
    const long long cell = g.lead + (long long)i0 * planestride + (long long)i1 * rowstride + i2;
    unsigned long long pin = (unsigned long long)(fin + cell);
    unsigned long long pout = (unsigned long long)(fout + cell);
    asm volatile("" : "+l"(pin), "+l"(pout));   // keep `pointer + constant-bank offset` as the address form
%(loads)s
%(body)s
%(stores)s"""

# Boundary entries folded into the fused kernel (TASKS instantiation).  Entry `f_k(c_out) = value` of a
# boundary method is read by exactly one pull: cell c_out + v_k, population k.  Instead of a list kernel
# storing the value into the ghost / solid cell and this kernel pulling it back, the block that owns the
# pulling cell evaluates its entries here -- one entry per thread, in parallel, same explicit
# round-to-nearest arithmetic as the list kernel k_bc (runtime) -- and hands the values to the owning
# threads through shared memory.  The host (boundary.plan_tasks) proves that no entry reads what another
# entry stores, so the input array alone determines every value.
#
# A thread computes CPT cells (1 for fp64 populations, 2 for fp32 populations: half the bytes per load,
# so twice the loads must be in flight to keep HBM busy at the same occupancy).  Coordinates and the
# loads of ALL its cells come first, then the cells are computed one after the other.
_CELL_COORDS = r"""    // ---- cell %(c)d of this thread
    const unsigned by%(c)d_ = offs.pair ? by * %(cpt)du + %(c)du : by;
    const unsigned bz%(c)d_ = offs.pair ? bz : bz * %(cpt)du + %(c)du;
    const int i1_%(c)d = g.lo[1] + (int)(by%(c)d_ * ty + (tid >> txshift));
    const int i0_%(c)d = g.lo[0] + (int)bz%(c)d_;
    const bool act%(c)d_ = !(i2 >= g.hi[2] || i1_%(c)d >= g.hi[1] || i0_%(c)d >= g.hi[0]);
    const long long cell%(c)d_ = g.lead + (long long)i0_%(c)d * planestride + (long long)i1_%(c)d * rowstride + i2;
    // the task range of the block is requested first and consumed after the population loads have
    // been issued, so that the memory round trips overlap
    int ta%(c)d_ = 0, tb%(c)d_ = 0;
    if (TASKS && i0_%(c)d < g.hi[0] && by%(c)d_ < offs.vgy) {
        const long long bid_ = ((long long)(i0_%(c)d - g.w[0]) * tasks.ngroups_y + by%(c)d_) * tasks.ngroups_x + blockIdx.x;
        ta%(c)d_ = __ldg(tasks.block_ptr + bid_);
        tb%(c)d_ = __ldg(tasks.block_ptr + bid_ + 1);
    }"""

_CELL_OPEN = r"""    {   // ======== cell %(c)d ========
    const int i0 = i0_%(c)d, i1 = i1_%(c)d;
    const bool active_ = act%(c)d_;
    const long long cell = cell%(c)d_;
    const int t0_ = ta%(c)d_, t1_ = tb%(c)d_;
    unsigned long long pout = (unsigned long long)(fout + cell);
    asm volatile("" : "+l"(pout));
    (void)i0; (void)i1;"""

_SMEM_DECL = r"""    extern __shared__ __align__(16) unsigned char lbmk_smem_[];
    real_c* const sm_val_ = (real_c*)lbmk_smem_;                                  // [NQ][LBMK_BLOCK]
    unsigned long long* const sm_mask_ = (unsigned long long*)(sm_val_ + NQ_ * LBMK_BLOCK);"""

_TASKS_CHAIN = r"""    unsigned long long tmask_ = 0ull;
    if (TASKS && t1_ > t0_) {                              // block-uniform
        sm_mask_[tid] = 0ull;
        __syncthreads();
        for (int t = t0_ + (int)tid; t < t1_; t += LBMK_BLOCK) {
            const unsigned code = __ldg(tasks.code + t);
            const long long l0_ = __ldg(tasks.l0 + t), l1_ = __ldg(tasks.l1 + t);
            const double* const rp_ = tasks.rhs[t];
            const double d = __ldg(tasks.dist + t);
            const unsigned kind = code >> 16, kk = (code >> 8) & 255u, th = code & 255u;
            const double a = (double)__ldg(fin + l0_);
            const double b = (double)__ldg(fin + l1_);     // (position 0 when the kind has one load)
            const double r = rp_ ? __ldg(rp_) : 0.0;
            double v = a;                                                      // Neumann
            if (kind == 0u) v = __dadd_rn(a, r);                               // bounce-back
            else if (kind == 1u) v = __dadd_rn(-a, r);                         // anti-bounce-back
            else if (kind != 4u) {
                const double far = __dmul_rn(__dsub_rn(1.0, d), b);
                const double near = __dmul_rn(d, a);
                v = __dadd_rn(__dadd_rn(far, kind == 2u ? near : -near), r);   // Bouzidi (anti-)bounce-back
            }
            sm_val_[kk * LBMK_BLOCK + th] = (real_c)(%(tin)s)v;           // rounded like the stored value
            atomicOr(sm_mask_ + th, 1ull << kk);
        }
        __syncthreads();
        tmask_ = sm_mask_[tid];
    }
    if (active_) {"""


def _inline_image(v, k, slab, tout):
    """store suffix for the image that only crosses the FASTEST axis (two lanes of every row): done
    with the main store from the register value -- no re-read, no divergent tail.  In the WALLZ
    instantiation the same lanes store the bounced value into the wall's ghost cell instead
    (bounce_back / anti_bounce_back of the NEXT step, reference boundary.py:462-464, 678-680, with the
    arithmetic of the list kernel k_bc: explicit round-to-nearest add, no contraction)."""
    aa_terms = ["aa%s%d_%d" % ("P" if v[a] > 0 else "M", a, abs(v[a])) for a in range(3) if v[a] != 0]
    aa = ""
    if aa_terms:
        # in-place even step: a population that leaves the interior is also stored at the fully wrapped
        # position (its periodic image); the unwrapped store above is what the boundary kernels read
        aa = " if (AAEVEN) { const long long wr_ = %s; if (wr_) __stcg(p_ + wr_, o_); }" % " + ".join(aa_terms)
    if v[2] == 0:
        return aa
    cond = "p2" if v[2] > 0 else "m2"
    if slab == 2:
        image = " if (%s) qbase[%dLL * qps + cell + d2] = o_;" % (cond, k)
    else:
        image = " if (%s) __stcg(p_ + d2, o_);" % cond
    wall, neg = ("wlo", "walls.neg_lo") if v[2] < 0 else ("whi", "walls.neg_hi")
    bounce = (" if (%s) __stcg((%s*)(pout + offs.wall[%d]), (%s)__dadd_rn(%s ? -(double)o_ : (double)o_, walls.rhs[%d]));"
              % (wall, tout, k, tout, neg, k))
    # in-place even step with walls: the bounced value goes into the cell's OWN slot of population k --
    # where the odd step reads the population entering from the wall (the transformed position of the
    # list entry (sym k, c + v_k) is (k, c)); that slot belongs to the ghost cell c + v_k in this step
    own = (" if (%s) __stcg((%s*)(pout + offs.nat[%d]), (%s)__dadd_rn(%s ? -(double)o_ : (double)o_, walls.rhs[%d]));"
           % (wall, tout, k, tout, neg, k))
    return aa + " if (!AAEVEN) { if (!WALLZ) {%s } else {%s } } else if (WALLZ) {%s }" % (image, bounce, own)


def _images_code(velocities, tout, slab):
    """tail of the fused kernel: stores of the needed (population, image) pairs that cross axis 0
    or 1 (cells of whole boundary rows / planes: full warps, rare)."""
    import itertools

    lines = ["    if (!AAEVEN && (d0 | d1)) {"]
    dname = ["d0", "d1", "d2"]
    moving = [k for k, v in enumerate(velocities) if v[0] != 0 or v[1] != 0]
    # re-read this thread's own stores, all loads issued back to back (one memory latency)
    for k in moving:
        lines.append("        const %s v%d_ = fout[%dLL * g.pstride + cell];" % (tout, k, k))
    for k in moving:
        v = velocities[k]
        axes = [a for a in range(3) if v[a] != 0]
        conds = {a: ("p%d" % a if v[a] > 0 else "m%d" % a) for a in axes}
        for r in range(1, len(axes) + 1):
            for sub in itertools.combinations(axes, r):
                if sub == (2,):
                    continue   # done inline with the main store
                cond = " && ".join(conds[a] for a in sub)
                off = " + ".join(dname[a] for a in sub)
                if slab in sub:
                    target = "qbase[%dLL * qps + cell + %s]" % (k, off)
                else:
                    target = "fout[%dLL * g.pstride + cell + %s]" % (k, off)
                lines.append("        if (%s) %s = v%d_;" % (cond, target, k))
    lines.append("    }")
    return "\n".join(lines)


def _canonical(offset):
    offset = tuple(int(o) for o in offset)
    return (0,) * (3 - len(offset)) + offset


_LAUNCH_HEAD = 'extern "C" int lbmk_%(name)s(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream)'
_LAUNCH_HEAD_WALLS = ('static int lbmk_launch_%(name)s_(const void* fin, void* fout, const lbmk_grid* g, '
                      'const double* scalars, const lbmk_peers* peers, const lbmk_walls* walls, '
                      'const lbmk_tasks* tasks, void* stream, int aa_phase = 0)')
_LAUNCH_TAIL_AA = """
// in-place streaming: ONE array; phase 0 = even step (gather, scatter back), phase 1 = odd step (local)
extern "C" int lbmk_%(name)s_aa(void* f, const lbmk_grid* g, const double* scalars, int phase, void* stream)
{
    return lbmk_launch_%(name)s_(f, f, g, scalars, nullptr, nullptr, nullptr, stream, phase ? 2 : 1);
}
// the same with the bounce-back walls of the fastest axis applied by the kernel (walls may be NULL)
extern "C" int lbmk_%(name)s_aa_walls(void* f, const lbmk_grid* g, const double* scalars, int phase,
                                      const lbmk_walls* walls, void* stream)
{
    return lbmk_launch_%(name)s_(f, f, g, scalars, nullptr, walls, nullptr, stream, phase ? 2 : 1);
}
"""
_LAUNCH_TAIL_WALLS = """
extern "C" int lbmk_%(name)s_tasks(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                                   const lbmk_peers* peers, const lbmk_tasks* tasks, void* stream)
{
    if (!tasks) return lbmk_%(name)s_walls(fin, fout, g, scalars, peers, nullptr, stream);
    return lbmk_launch_%(name)s_(fin, fout, g, scalars, peers, nullptr, tasks, stream);
}
extern "C" int lbmk_%(name)s_peers(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                                   const lbmk_peers* peers, void* stream)
{
    return lbmk_%(name)s_walls(fin, fout, g, scalars, peers, nullptr, stream);
}
extern "C" int lbmk_%(name)s(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream)
{
    return lbmk_%(name)s_walls(fin, fout, g, scalars, nullptr, nullptr, stream);
}
"""
_LAUNCH_TAIL_WALLS = """
extern "C" int lbmk_%(name)s_walls(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                                   const lbmk_peers* peers, const lbmk_walls* walls, void* stream)
{
    return lbmk_launch_%(name)s_(fin, fout, g, scalars, peers, walls, nullptr, stream);
}""" + _LAUNCH_TAIL_WALLS
_CALL_PLAIN = """    lbmk_kernel_%(name)s<<<grid, LBMK_BLOCK, 0, (cudaStream_t)stream>>>(
        (const %(tin)s*)fin, (%(tout)s*)fout, *g, offs%(scalar_args)s);"""
_CALL_WALLS = """    const lbmk_peers pr_ = peers ? *peers : lbmk_peers{nullptr, nullptr, 0, 0, 0};
    const lbmk_tasks notasks_ = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0};
    cudaLaunchConfig_t cfg_;
    memset(&cfg_, 0, sizeof(cfg_));
    cfg_.gridDim = grid;
    cfg_.blockDim = dim3(LBMK_BLOCK);
    cfg_.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr_[1];
    attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr_[0].val.programmaticStreamSerializationAllowed = 1;
    cfg_.attrs = attr_;
    cfg_.numAttrs = (g->wrap & LBMK_WRAP_PDL) ? 1 : 0;
    lbmk_grid gk_ = *g;
    gk_.wrap &= 7;
    if (tasks) {
        // the table maps cells to (block, thread) of THIS launch geometry
        if (walls || tasks->tx != g->tx || g->lo[1] != g->w[1] || g->lo[2] != g->w[2] || offs.fold
            || tasks->ngroups_x != (int)grid.x || tasks->ngroups_y != (int)offs.vgy) return -4;
        const size_t smem_ = (size_t)%(nin)d * LBMK_BLOCK * sizeof(real_c_%(name)s) + LBMK_BLOCK * sizeof(unsigned long long);
        static bool attr_set_ = false;
        if (!attr_set_ && smem_ > 48 * 1024) {
            cudaFuncSetAttribute(lbmk_kernel_%(name)s<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_);
            attr_set_ = true;
        }
        cfg_.dynamicSmemBytes = smem_;
        cudaLaunchKernelEx(&cfg_, lbmk_kernel_%(name)s<0, true>,
            (const %(tin)s*)fin, (%(tout)s*)fout, gk_, offs, pr_, img, lbmk_walls{}, *tasks%(scalar_args)s);
    }%(aa_call)s else if (walls)
        cudaLaunchKernelEx(&cfg_, lbmk_kernel_%(name)s<1, false>,
            (const %(tin)s*)fin, (%(tout)s*)fout, gk_, offs, pr_, img, *walls, notasks_%(scalar_args)s);
    else
        cudaLaunchKernelEx(&cfg_, lbmk_kernel_%(name)s<0, false>,
            (const %(tin)s*)fin, (%(tout)s*)fout, gk_, offs, pr_, img, lbmk_walls{}, notasks_%(scalar_args)s);"""
_CALL_AA = """ else if (aa_phase == 1 && walls)
        cudaLaunchKernelEx(&cfg_, lbmk_kernel_%(name)s<3, false>,
            (const %(tin)s*)fin, (%(tout)s*)fout, gk_, offs, pr_, img, *walls, notasks_%(scalar_args)s);
    else if (aa_phase == 1)
        cudaLaunchKernelEx(&cfg_, lbmk_kernel_%(name)s<2, false>,
            (const %(tin)s*)fin, (%(tout)s*)fout, gk_, offs, pr_, img, lbmk_walls{}, notasks_%(scalar_args)s);"""


def kernel_source(ir, storage="double", cse=True, minblocks=1, images=False, slab=0, compute="double", cpt=1,
                  aa=False):
    """CUDA source of one per-cell kernel + its C-ABI launcher.  `compute` is the arithmetic type of
    the kernel body (double, or float for the all-fp32 mode of the fused kernel); `cpt` the number of
    cells a thread of the fused kernel computes."""
    temps, outs = lower_statements(ir.statements, ir.outputs, cse=cse)
    pr = (_explicit_printers if (aa and images) else _printers)[compute == "float"]
    nq = len(ir.in_syms)
    tin = "real_m" if ir.in_array == "m" else "real_f"
    tout = "real_m" if ir.out_array == "m" else "real_f"
    if not images:
        cpt = 1
    body = ["    const real_c %s = %s;" % (lhs, pr.doprint(rhs)) for lhs, rhs in temps]
    vels = [tuple(-o for o in _canonical(off)) for off in ir.in_offsets]
    if images:
        if aa:
            # a library for in-place streaming evaluates every output in the straight-line body, before any
            # store: the even-step, odd-step and two-array instantiations then contract the same a*b+c into
            # FMAs (with the expression inside the store blocks nvcc contracted 4 of the 19 D3Q19 outputs
            # differently in the even-step instantiation: last-bit differences), so they are bit-identical
            body = body + ["    const real_c o%d_v = %s;" % (k, pr.doprint(o)) for k, o in enumerate(outs)]
            out_text = ["o%d_v" % k for k in range(len(outs))]
        else:
            out_text = [pr.doprint(o) for o in outs]
        stores = [
            "    { const %s o_ = (%s)(%s); %s* p_ = (%s*)(pout + offs.out[%d]); __stcg(p_, o_);%s }"
            % (tout, tout, out_text[k], tout, tout, k, _inline_image(vels[k], k, slab, tout))
            for k, o in enumerate(outs)
        ]
        parts = [_SMEM_DECL]
        for c in range(cpt):
            parts.append(_CELL_COORDS % dict(c=c, cpt=cpt))
        # a cell beyond the range has every later cell of the thread beyond it too
        parts.append("    if (!TASKS && !act0_) return;")
        # raw loads of all cells first (no conversion: nothing consumes them before the task chains and
        # the loads of the other cells have been issued)
        for c in range(cpt):
            parts.append("    %s %s;" % (tin, ", ".join("r%d_%d = 0" % (c, k) for k in range(nq))))
            parts.append("    if (act%d_) {" % c)
            parts.append("        unsigned long long pin = (unsigned long long)(fin + cell%d_);" % c)
            parts.append('        asm volatile("" : "+l"(pin));   // keep `pointer + constant-bank offset` as the address form')
            for k in range(nq):
                parts.append("        r%d_%d = LBMK_LDG((const %s*)(pin + offs.in[%d]));" % (c, k, tin, k))
            parts.append("    }")
        images_tail = _images_code(vels, tout, slab)
        for c in range(cpt):
            parts.append(_CELL_OPEN % dict(c=c))
            parts.append(_TASKS_CHAIN % dict(tin=tin))
            import os

            intconv = os.environ.get("PYLBM_B200_F2D") == "int" and compute == "double"
            for k, sym in enumerate(ir.in_syms):
                parts.append("    real_c %s = %sr%d_%d%s;" % (sym, "lbmk_f2d(" if intconv else "(real_c)", c, k,
                                                              ")" if intconv else ""))
            parts.append("    if (TASKS && tmask_) {   // pulled values that are boundary entries of this step")
            for k, sym in enumerate(ir.in_syms):
                parts.append("        if (tmask_ & (1ull << %d)) %s = sm_val_[%d * LBMK_BLOCK + tid];" % (k, sym, k))
            parts.append("    }")
            shifts = []
            for a in range(3):
                for mag in sorted({abs(v[a]) for v in vels if v[a] != 0}):
                    on = "AAEVEN && !WALLZ" if a == 2 else "AAEVEN"   # walls of the fastest axis: no image
                    shifts.append("    const long long aaP%d_%d = (%s && i%d + %d >= g.n[%d] - g.w[%d]) ? img.dhigh[%d] : 0LL;"
                                  % (a, mag, on, a, mag, a, a, a))
                    shifts.append("    const long long aaM%d_%d = (%s && i%d - %d < g.w[%d]) ? img.dlow[%d] : 0LL;"
                                  % (a, mag, on, a, mag, a, a))
                    shifts.append("    (void)aaP%d_%d; (void)aaM%d_%d;" % (a, mag, a, mag))
            parts.append(_IMAGES_PROLOGUE % dict(tout=tout, aa_shifts="\n".join(shifts)))
            parts.extend(body)
            parts.extend(stores)
            parts.append(images_tail)
            parts.append("    }   // active_")
            parts.append("    }   // cell %d" % c)
        thread_body = "\n".join(parts)
    else:
        loads = ["    const real_c %s = (real_c)__ldg((const %s*)(pin + offs.in[%d]));" % (sym, tin, k)
                 for k, sym in enumerate(ir.in_syms)]
        stores = [
            "    __stcg((%s*)(pout + offs.out[%d]), (%s)(%s));" % (tout, k, tout, pr.doprint(o))
            for k, o in enumerate(outs)
        ]
        thread_body = _THREAD_PLAIN % dict(loads="\n".join(loads), body="\n".join(body), stores="\n".join(stores))
    add, mul, div = count_ops(temps, outs)
    scal_params = "".join(", const real_c_%s %s" % (ir.name, _c_name(s)) for s in ir.scalars)
    scal_args = "".join(", (real_c_%s)scalars[%d]" % (ir.name, i) for i in range(len(ir.scalars)))
    table = ", ".join("{%d, %d, %d}" % _canonical(off) for off in ir.in_offsets)
    symmetric = getattr(ir, "symmetric", None) or list(range(len(outs)))
    sym_table = ", ".join(str(int(k)) for k in symmetric)
    src = "typedef %s real_c_%s;\n" % (compute, ir.name) + _KERNEL % dict(
        name=ir.name,
        tin=tin,
        tout=tout,
        tc=compute,
        nin=nq,
        nout=len(outs),
        cpt=cpt,
        ops="%d add, %d mul, %d div before FMA contraction, %s arithmetic, %d cell(s) per thread"
            % (add, mul, div, compute, cpt),
        minblocks=minblocks,
        scalar_params=scal_params,
        scalar_args=scal_args,
        offset_table=table,
        thread_body=thread_body,
        template="template <int MODE, bool TASKS>\n" if images else "",
        mode_decl=("    constexpr bool WALLZ = (MODE == 1 || MODE == 3), AAEVEN = (MODE == 2 || MODE == 3); "
                   "(void)WALLZ; (void)AAEVEN;\n" if images else ""),
        restrict="LBMK_RESTRICT" if images else "__restrict__",
        inpop_table=", ".join(str(int(k)) for k in (getattr(ir, "in_pops", None) or range(nq))),
        peer_param=(", const lbmk_peers pr, const lbmk_images img, const lbmk_walls walls, const lbmk_tasks tasks"
                    if images else ""),
        images_launch=(_IMAGES_LAUNCH % dict(nout=len(outs), tout=tout, sym_table=sym_table)) if images else "",
        kernel_call=(_CALL_WALLS if images else _CALL_PLAIN) % dict(
            name=ir.name, tin=tin, tout=tout, scalar_args=scal_args, nin=nq,
            aa_call=(_CALL_AA % dict(name=ir.name, tin=tin, tout=tout, scalar_args=scal_args)) if (images and aa) else ""),
        launch_head=(_LAUNCH_HEAD_WALLS if images else _LAUNCH_HEAD) % dict(name=ir.name),
        launch_tail=((_LAUNCH_TAIL_WALLS + (_LAUNCH_TAIL_AA if aa else "")) % dict(name=ir.name)) if images else "",
    )
    # user symbols may not be valid C identifiers (e.g. `lambda`)
    for s_ in ir.scalars:
        if _c_name(s_) != s_:
            src = _rename_identifier(src, s_, _c_name(s_))
    return src, (add, mul, div)


_C_KEYWORDS = {"tasks", "lambda", "double", "int", "float", "long", "short", "register", "const", "void", "auto", "g", "fin", "fout", "cell", "tid",
               "pin", "pout", "offs", "pr", "tx", "ty", "i0", "i1", "i2", "d0", "d1", "d2", "real_c"}


def _c_name(name):
    safe = "".join(ch if (ch.isalnum() or ch == "_") else "_" for ch in name)
    if safe in _C_KEYWORDS or safe[0].isdigit():
        safe = safe + "_"
    return safe


def _rename_identifier(src, old, new):
    import re

    return re.sub(r"(?<![A-Za-z0-9_])%s(?![A-Za-z0-9_])" % re.escape(old), new, src)


_DESCRIBE = r"""
extern "C" int lbmk_abi_version(void) { return LBMK_ABI_VERSION; }

extern "C" const char* lbmk_describe(void)
{
    return %(json)s;
}
"""


def default_cpt(storage):
    """cells per thread of the fused kernel.  1: measured on B200 (D3Q19 512^3, profiles/r02_fp32_modes.md) two
    cells per thread -- twice the loads in flight per thread, but 128 registers, 16 resident warps -- is
    SLOWER with fp32 populations and fp64 arithmetic (0.62 vs 0.77 of the roofline: that kernel is bound by
    warp-level parallelism for its fp64 / conversion chains, not by bytes in flight) and within noise for
    the all-fp32 kernel (0.924 vs 0.916).  PYLBM_B200_CPT=2 builds the two-cell variant."""
    import os

    env = os.environ.get("PYLBM_B200_CPT")
    if env:
        return max(1, min(4, int(env)))
    return 1


def default_minblocks(nv, compute="double", cpt=1):
    """resident 128-thread blocks per SM requested through __launch_bounds__ for the fused
    kernel (caps registers: 65536 / (128 * minblocks)); tuned on B200, see DESIGN.md:
    fp64 arithmetic 5 (<= 102 registers; 6 spills), fp32 arithmetic 8 (56 registers; 5 -> 8 gave
    3.53 -> 3.41 ms on the 512^3 D3Q19 kernel, 10 the same)."""
    import os

    env = os.environ.get("PYLBM_B200_MINBLOCKS")
    if env:
        return int(env)
    if cpt > 1:          # the raw loads of the other cells stay live while one cell is computed
        return 6 if compute == "float" else 4
    if compute == "float":
        return 8
    # fp64: D3Q27 (27 x 2 registers of populations alone) spills at 5 blocks; measured on the 512x256x256
    # channel: (128, 4) 2.385 ms, (128, 5) 2.408 ms, (128, 6) 2.447 ms per step
    return 4 if nv >= 24 else 5


def kernel_tag(kernels, dim, nv, storage="double", cse=True, compute="double", aa=False):
    """
    Cache key of a kernel library: hash of the IR (deterministic across processes, unlike the text
    produced by sympy.cse whose temporaries depend on set ordering), of the generator options and
    of this generator's own source.
    """
    h = hashlib.sha256()
    with open(__file__.replace(".pyc", ".py"), "rb") as fh:
        h.update(fh.read())
    import os

    cpt = default_cpt(storage)
    h.update(repr((ABI_VERSION, dim, nv, storage, cse, default_minblocks(nv, compute, cpt), compute, cpt,
                   os.environ.get("PYLBM_B200_F2D", ""), bool(aa))).encode())
    for ir in kernels:
        h.update(repr((ir.name, ir.in_array, ir.out_array, bool(ir.inner), list(ir.scalars),
                       [str(s) for s in ir.in_syms], [tuple(o) for o in ir.in_offsets],
                       list(getattr(ir, "in_pops", None) or []))).encode())
        for lhs, rhs in ir.statements:
            h.update((str(lhs) + "=" + sp.srepr(sp.sympify(rhs))).encode())
        for out in ir.outputs:
            h.update(sp.srepr(sp.sympify(out)).encode())
    return h.hexdigest()[:20]


def generate_source(kernels, dim, nv, storage="double", cse=True, compute="double", aa=False):
    """
    Full translation unit for a list of KernelIR.  Returns (source, info dict).
    `storage` is the type of the populations in HBM, `compute` the arithmetic type of the
    time-step kernels (one_time_step, transport); the whole-array kernels that produce or consume
    moments (f2m, m2f, equilibrium, ...) always compute in double.
    """
    if compute == "float" and storage != "float":
        raise ValueError("float arithmetic needs float storage of the populations")
    import json

    parts = [_HEADER % dict(abi=ABI_VERSION, storage=storage, slab=3 - dim, aa=1 if aa else 0)]
    info = {"abi": ABI_VERSION, "dim": dim, "nv": nv, "storage": storage, "compute": compute, "routines": {}}
    cpt = default_cpt(storage)
    for ir in kernels:
        fused = ir.name == "one_time_step"
        src, ops = kernel_source(ir, storage=storage, cse=cse, images=fused,
                                 minblocks=default_minblocks(nv, compute, cpt) if fused else 1, slab=3 - dim,
                                 compute=compute if ir.name in ("one_time_step", "transport") else "double",
                                 cpt=cpt, aa=aa)
        parts.append(src)
        info["routines"][ir.name] = {
            "scalars": list(ir.scalars),
            "in": ir.in_array,
            "out": ir.out_array,
            "inner": bool(ir.inner),
            "ops": {"add": ops[0], "mul": ops[1], "div": ops[2]},
        }
    text = json.dumps(info, sort_keys=True)
    literal = '"' + text.replace("\\", "\\\\").replace('"', '\\"') + '"'
    parts.append(_DESCRIBE % dict(json=literal))
    source = "\n".join(parts)
    info["hash"] = kernel_tag(kernels, dim, nv, storage, cse, compute, aa=aa)
    info["aa"] = bool(aa)
    return source, info
