"""
Per-cell kernels of the pull algorithm as a small sequential IR.

What is computed follows the reference's symbolic algorithm
(reference: pylbm/algorithm/pull.py:11-59 fused pull step,
pylbm/algorithm/base.py:298-334 f2m, 346-369 m2f, 381-393 equilibrium,
404-428 relaxation, 255-263 restore conserved moments, 439-495 + ode.py:11-16
half-step explicit-Euler source terms).  Matrix equations are element-wise
*sequential in-place* assignments in the reference's generated code
(pylbm/generator/printing/cython.py:316-350), which is kept: a later row sees
the already-updated earlier rows.

How it is computed is chosen for the GPU: with a relative velocity the shifted
moments are evaluated as T(u)·(M f) and the populations as M^{-1}·(T(-u) m)
(numeric Q×Q transforms in registers + a sparse polynomial shift), instead of
the reference's dense polynomial matrices (T(u) M) and (M^{-1} T(-u)).

A KernelIR is a list of `(symbol, expr)` statements over the input symbols
`f0..` (or `m0..`), runtime scalars (t, dt, user parameters) and locals; the
CUDA lowering (cudagen.py) turns it into SSA, applies CSE and prints it.
"""

import sympy as sp

from .scheme import rel_ux, rel_uy, rel_uz

__all__ = ["KernelIR", "PullAlgorithm"]

DT = sp.Symbol("dt")
T = sp.Symbol("t")


class KernelIR:
    """
    One per-cell map.

    name        routine name (one_time_step, f2m, ...)
    in_array    'f' or 'm': array the inputs are read from
    in_syms     Q input symbols; in_offsets[k] = integer offset (per dimension)
                added to the cell index when reading population k
    out_array   'fnew', 'f' or 'm'
    outputs     Q expressions/symbols stored at the cell, population by population
    statements  sequential (symbol, expr) assignments evaluated before the stores
    inner       True: loop on [vmax, n-vmax) per axis; False: whole array
    scalars     names of runtime scalar arguments, in call order
    """

    def __init__(self, name, in_array, in_syms, in_offsets, out_array, statements, outputs, inner, scalars):
        self.name = name
        self.in_array = in_array
        self.in_syms = in_syms
        self.in_offsets = in_offsets
        self.out_array = out_array
        self.statements = statements
        self.outputs = outputs
        self.inner = inner
        self.scalars = scalars


def _horner(expr, syms):
    """polynomial in the relative-velocity symbols -> nested (Horner) form: fewer multiplications
    than the fully expanded products (D3Q27 central moments: 2323 -> 1246 operations)."""
    expr = sp.expand(expr)
    used = [s for s in syms if expr.has(s)]
    if not used:
        return expr
    try:
        return sp.horner(expr, *used)
    except Exception:
        return expr


def _recursive_sub(expr, replace):
    for _ in range(len(replace) + 1):
        new = expr.subs(replace)
        if new == expr:
            return new
        expr = new
    return expr


class PullAlgorithm:
    """
    Builds the kernels of a scheme.  `settings` accepts the reference's keys
    (m_local, split, check_isfluid).  `m_local` and `split` only change how the
    reference arranges its loops (one loop per statement over a global m array,
    base.py:575-593), not what a cell computes: the fused kernel is generated
    either way.  `check_isfluid` is refused (the reference's Cython backend
    cannot print its `If` either, printing/cython.py:233-244).
    """

    def __init__(self, scheme, settings=None):
        self.scheme = scheme
        self.dim = scheme.dim
        self.ns = int(scheme.stencil.nv_ptr[-1])
        self.settings = settings or {}
        if self.settings.get("check_isfluid", False):
            raise NotImplementedError("check_isfluid is not supported by the CUDA backend")
        self.nconsm = len(scheme.consm)
        self.velocities = scheme.stencil.get_all_velocities()

        ns = self.ns
        self.m = [sp.Symbol("m%d" % i, real=True) for i in range(ns)]
        self.f = [sp.Symbol("f%d" % i, real=True) for i in range(ns)]

        params = [(sp.Symbol(str(k)), v) for k, v in scheme.param.items()]
        params += list(scheme.param.items())
        moments = [(k, self.m[int(i)]) for k, i in scheme.consm.items()]
        self._subs_params = params
        self._subs_full = params + moments

        self.M = scheme.M
        self.invM = scheme.invM
        self.eq = sp.Matrix([_recursive_sub(sp.sympify(e), self._subs_full) for e in scheme.EQ])
        self.s = sp.Matrix([_recursive_sub(sp.sympify(e), self._subs_full) for e in scheme.s])

        self.with_rel_vel = scheme.rel_vel is not None
        if self.with_rel_vel:
            self.rel_sym = [rel_ux, rel_uy, rel_uz][: self.dim]
            self.rel_vel = [_recursive_sub(sp.sympify(e), self._subs_full) for e in scheme.rel_vel]
            self.Tu = scheme.Tu
            self.Tmu = scheme.Tmu

        self.source_eq = []
        for source in scheme._source_terms:
            if source:
                for k, v in source.items():
                    lhs = _recursive_sub(sp.sympify(k), self._subs_full)
                    rhs = _recursive_sub(sp.sympify(v), self._subs_full)
                    if lhs not in self.m:
                        raise ValueError("source term on %s: not a conserved moment" % k)
                    self.source_eq.append((lhs, rhs))

        coords = set(sp.Symbol(str(c)) for c in scheme.symb_coord) | set(scheme.symb_coord)
        for expr in list(self.eq) + list(self.s) + [r for _, r in self.source_eq]:
            if expr.free_symbols & coords:
                raise NotImplementedError(
                    "space-dependent equilibrium / relaxation / source terms are not supported"
                )

    # ------------------------------------------------------------------
    def _scalars(self, statements, outputs, inputs):
        known = set(inputs)
        for lhs, _ in statements:
            known.add(lhs)
        free = set()
        for _, rhs in statements:
            free |= sp.sympify(rhs).free_symbols
        for out in outputs:
            free |= sp.sympify(out).free_symbols
        names = sorted(str(s) for s in free - known)
        return names

    def _ir(self, name, in_array, in_syms, in_offsets, out_array, statements, outputs, inner):
        scalars = self._scalars(statements, outputs, in_syms)
        return KernelIR(name, in_array, in_syms, in_offsets, out_array, statements, outputs, inner, scalars)

    def _zero_offsets(self):
        return [(0,) * self.dim] * self.ns

    # ---- local pieces -------------------------------------------------
    def _linear(self, mat, vec):
        return [sum((mat[i, j] * vec[j] for j in range(self.ns)), sp.Integer(0)) for i in range(self.ns)]

    def _f2m_local(self, f):
        """statements computing m from the (already pulled) populations f."""
        m, nc = self.m, self.nconsm
        raw = self._linear(self.M, f)
        if not self.with_rel_vel:
            return [(m[i], raw[i]) for i in range(self.ns)], None
        mraw = [sp.Symbol("mraw%d" % i, real=True) for i in range(self.ns)]
        stm = [(mraw[i], raw[i]) for i in range(self.ns)]
        stm += [(m[i], mraw[i]) for i in range(nc)]
        stm += [(self.rel_sym[d], self.rel_vel[d]) for d in range(self.dim)]
        shifted = self.Tu * sp.Matrix(mraw)
        stm += [(m[i], _horner(shifted[i], self.rel_sym)) for i in range(nc, self.ns)]
        return stm, mraw

    def _source_local(self):
        return [(lhs, lhs + DT / 2 * rhs) for lhs, rhs in self.source_eq]

    def _relaxation_local(self):
        eq = self.eq
        if self.with_rel_vel:
            eq = self.Tu * eq
            if not self.source_eq:
                # rel_u still equals its definition in terms of the (unchanged) conserved moments:
                # substituting it collapses the shifted equilibria (central-moment equilibria are
                # constants times the density), which the reference evaluates as long polynomials
                rel = dict(zip(self.rel_sym, self.rel_vel))
                eq = eq.applyfunc(lambda e: sp.factor(sp.cancel(sp.together(sp.expand(e.subs(rel))))))
            else:
                eq = eq.applyfunc(lambda e: _horner(e, self.rel_sym))
        stm = []
        for i in range(self.ns):
            if self.s[i] == 0:
                continue  # m_i = m_i
            stm.append((self.m[i], (1 - self.s[i]) * self.m[i] + self.s[i] * eq[i]))
        return stm

    def _m2f_local(self):
        m = self.m
        if self.with_rel_vel:
            back = [sp.Symbol("mback%d" % i, real=True) for i in range(self.ns)]
            shifted = self.Tmu * sp.Matrix(m)
            stm = [(back[i], _horner(shifted[i], self.rel_sym)) for i in range(self.ns)]
            return stm, self._linear(self.invM, back)
        return [], self._linear(self.invM, m)

    # ---- kernels ------------------------------------------------------
    def one_time_step(self):
        """fused pull stream + collide (reference: algorithm/pull.py:11-59)."""
        f = self.f
        stm, mraw = self._f2m_local(f)
        stm += self._source_local()
        stm += self._relaxation_local()
        if self.with_rel_vel:
            shifted = self.Tu * sp.Matrix(mraw)
            if not self.source_eq:
                # same remark: e.g. the shifted momentum q - u*rho is identically zero
                rel = dict(zip(self.rel_sym, [e.subs(dict(zip(self.m, mraw))) for e in self.rel_vel]))
                rows = [sp.factor(sp.cancel(sp.together(sp.expand(shifted[i].subs(rel))))) for i in range(self.nconsm)]
            else:
                rows = [_horner(shifted[i], self.rel_sym) for i in range(self.nconsm)]
            stm += [(self.m[i], rows[i]) for i in range(self.nconsm)]
        stm += self._source_local()
        tail, outputs = self._m2f_local()
        stm += tail
        offsets = [tuple(-int(c) for c in v) for v in self.velocities]
        ir = self._ir("one_time_step", "f", f, offsets, "fnew", stm, outputs, True)
        # population with the opposite velocity inside the same sub-scheme (bounce-back partner)
        ir.symmetric = [int(k) for k in self.scheme.stencil.get_symmetric()]
        return ir

    def transport(self):
        offsets = [tuple(-int(c) for c in v) for v in self.velocities]
        return self._ir("transport", "f", self.f, offsets, "fnew", [], list(self.f), True)

    def f2m(self):
        """m = M f on the whole array (no relative velocity: base.py:336-344)."""
        out = self._linear(self.M, self.f)
        return self._ir("f2m", "f", self.f, self._zero_offsets(), "m", [], out, False)

    def f2m_consm(self):
        """conserved moments only (rows 0..nconsm-1 of M f): what `sol.m[symbol]` reads back after a
        step, without materialising the other Q - nconsm rows (reference: simulation.py:215-224 runs the
        full f2m over the whole array)."""
        out = self._linear(self.M, self.f)[: self.nconsm]
        return self._ir("f2m_consm", "f", self.f, self._zero_offsets(), "m", [], out, False)

    def m2f(self):
        out = self._linear(self.invM, self.m)
        return self._ir("m2f", "m", self.m, self._zero_offsets(), "f", [], out, False)

    def equilibrium(self):
        stm = [(self.m[i], self.eq[i]) for i in range(self.ns) if self.eq[i] != self.m[i]]
        return self._ir("equilibrium", "m", self.m, self._zero_offsets(), "m", stm, list(self.m), False)

    def relaxation(self):
        eq = self.eq
        stm = [
            (self.m[i], (1 - self.s[i]) * self.m[i] + self.s[i] * eq[i])
            for i in range(self.ns)
            if self.s[i] != 0
        ]
        return self._ir("relaxation", "m", self.m, self._zero_offsets(), "m", stm, list(self.m), False)

    def source_term(self):
        stm = self._source_local()
        return self._ir("source_term", "m", self.m, self._zero_offsets(), "m", stm, list(self.m), True)

    def _swapped(self, ir, name):
        """the same map reading an in-place (AA) array after an even step: population k of cell x sits in
        slot (kbar, x + v_k) (include/lbm_b200.h: lbm_sim_set_aa); interior cells only."""
        sym = [int(k) for k in self.scheme.stencil.get_symmetric()]
        offsets = [tuple(int(c) for c in v) for v in self.velocities]
        out = self._ir(name, ir.in_array, ir.in_syms, offsets, ir.out_array, ir.statements, ir.outputs, True)
        out.in_pops = sym
        return out

    def kernels(self, aa=False):
        out = [self.transport(), self.f2m(), self.f2m_consm(), self.m2f(), self.relaxation(), self.equilibrium(),
               self.one_time_step()]
        if self.source_eq:
            out.append(self.source_term())
        if aa:
            out += [self._swapped(self.f2m(), "f2m_sw"), self._swapped(self.f2m_consm(), "f2m_consm_sw")]
        return out
