"""
d'Humieres-type lattice Boltzmann scheme: moment matrix M, its inverse, the
relative-velocity shift T(u), equilibria, relaxation rates, conserved moments.

Mirror of pylbm.scheme.Scheme for what the time-step path consumes
(reference: pylbm/scheme.py:136-190 constructor, 202-223 equilibrium,
225-241 conserved moments in front, 464-532 moment matrices, 575-607
conserved-moment lookup).  Host-side symbolic setup, runs once.

Differences in *how* (not in what) it is computed: the user parameters are
substituted before the matrices are built, and numbers are turned into exact
rationals, so M is inverted exactly and T(u) = M(u) M^{-1} is a polynomial
matrix with rational coefficients (the reference inverts a symbolic matrix and
substitutes afterwards; the values agree to rounding).
"""

from collections import OrderedDict

import sympy as sp

from .stencil import Stencil

__all__ = ["Scheme", "rel_ux", "rel_uy", "rel_uz"]

rel_ux, rel_uy, rel_uz = sp.symbols("rel_ux, rel_uy, rel_uz", real=True)


def _rationalize(expr):
    """Floats -> exact rationals (same intent as the reference's nsimplify)."""
    expr = sp.sympify(expr)
    if not expr.atoms(sp.Float):
        return expr
    return sp.nsimplify(expr, rational=True)


class Scheme:
    """
    Attributes mirrored from the reference: `dim`, `stencil`, `la`, `param`,
    `rel_vel`, `symb_t`, `symb_coord`, `nschemes`, `P`, `s`, `EQ`, `M`,
    `invM`, `Tu`, `Tmu`, `consm` (symbol -> row, conserved rows first),
    `_source_terms`, and the `*_no_swap` copies.
    M, invM, Tu, Tmu already have `param` substituted.
    """

    def __init__(self, dico, check_inverse=False, need_validation=True):
        self.stencil = Stencil(dico, need_validation=False)
        self.dim = self.stencil.dim
        self.param = dico.get("parameters", {}) or {}
        self.la = dico["scheme_velocity"]
        self.rel_vel = dico.get("relative_velocity", None)
        self.symb_t = self.param.get("time", sp.Symbol("t"))
        self.symb_coord = [
            self.param.get("space_x", sp.Symbol("X")),
            self.param.get("space_y", sp.Symbol("Y")),
            self.param.get("space_z", sp.Symbol("Z")),
        ]
        self.nschemes = self.stencil.nstencils
        schemes = dico["schemes"]
        ns = int(self.stencil.nv_ptr[-1])

        self._check_entry_size(schemes, "relaxation_parameters")
        self.s = sp.Matrix([r for s in schemes for r in s["relaxation_parameters"]])

        if len(schemes) == 1 and "M" in schemes[0]:
            self.P = []
            self.M = sp.Matrix(schemes[0]["M"]).subs(list(self.param.items()))
            self.invM = self.M.inv()
            self.Tu = sp.eye(ns)
            self.Tmu = sp.eye(ns)
        else:
            self._check_entry_size(schemes, "polynomials")
            self.P = sp.Matrix([p for s in schemes for p in s["polynomials"]])
            self.M, self.invM, self.Tu, self.Tmu = self._moment_matrices()

        self._source_terms = [s.get("source_terms", None) for s in schemes]
        self.EQ = self._equilibrium(schemes)

        self.s_no_swap = self.s.copy()
        self.EQ_no_swap = self.EQ.copy()
        self.M_no_swap = self.M.copy()
        self.invM_no_swap = self.invM.copy()

        self.consm = self._conserved_moments(schemes)
        self._conserved_in_front()

        if check_inverse:
            if sp.simplify(self.M * self.invM - sp.eye(ns)) != sp.zeros(ns, ns):
                raise ValueError("M * invM is not the identity")

    # ------------------------------------------------------------------
    def _check_entry_size(self, schemes, key):
        for i, s in enumerate(schemes):
            if len(s[key]) != self.stencil.nv[i]:
                raise ValueError(
                    "the size of the entry for the key {0} in the scheme {1} has not the same "
                    "size of the stencil {1}: {2}, {3}".format(key, i, len(s[key]), self.stencil.nv[i])
                )

    def _moment_matrices(self):
        """
        M[i, j] = P_i(la v_j); with a relative velocity, M(u)[i, j] = P_i(la v_j - u)
        and T(u) = M(u) M^{-1}, block by block (reference: pylbm/scheme.py:464-532).
        """
        ns = int(self.stencil.nv_ptr[-1])
        M = sp.zeros(ns, ns)
        invM = sp.zeros(ns, ns)
        Tu = sp.eye(ns)
        params = list(self.param.items())
        la = _rationalize(sp.sympify(self.la).subs(params))
        u_tilde = [rel_ux, rel_uy, rel_uz]
        coords = [sp.Symbol(str(c)) for c in self.symb_coord]

        for k, vel in enumerate(self.stencil.v):
            lo, hi = int(self.stencil.nv_ptr[k]), int(self.stencil.nv_ptr[k + 1])
            nvk = hi - lo
            polys = []
            for i in range(lo, hi):
                p = sp.sympify(self.P[i])
                # symbols are matched by name, as the reference does (subs with str keys)
                p = p.subs([(s, sp.Symbol(s.name)) for s in p.free_symbols])
                p = p.subs([(sp.Symbol(str(key)), val) for key, val in params])
                polys.append(sp.expand(_rationalize(p)))
            Mk = sp.zeros(nvk, nvk)
            for j, vj in enumerate(vel):
                point = {coords[d]: sp.Integer(vj.v[d]) * la for d in range(self.dim)}
                for i in range(nvk):
                    Mk[i, j] = polys[i].subs(point)
            if Mk.free_symbols:
                raise ValueError(
                    "the moment matrix still depends on {}: check the 'parameters' entry".format(
                        Mk.free_symbols
                    )
                )
            invMk = Mk.inv()
            M[lo:hi, lo:hi] = Mk
            invM[lo:hi, lo:hi] = invMk
            if self.rel_vel is not None:
                Muk = sp.zeros(nvk, nvk)
                for j, vj in enumerate(vel):
                    point = {
                        coords[d]: sp.Integer(vj.v[d]) * la - u_tilde[d] for d in range(self.dim)
                    }
                    for i in range(nvk):
                        Muk[i, j] = polys[i].subs(point)
                Tu[lo:hi, lo:hi] = (Muk * invMk).applyfunc(sp.expand)
        Tmu = Tu.subs([(u, -u) for u in u_tilde], simultaneous=True)
        return M, invM, Tu, Tmu

    def _equilibrium(self, schemes):
        eq = []
        for i, s in enumerate(schemes):
            feq = s.get("feq", None)
            meq = s.get("equilibrium", None)
            if feq and meq:
                raise ValueError(
                    "Error in the creation of the scheme %d: you can have only 'feq' or 'equilibrium'" % i
                )
            if meq:
                eq.extend(sp.sympify(e) for e in meq)
            if feq:
                sli = slice(int(self.stencil.nv_ptr[i]), int(self.stencil.nv_ptr[i + 1]))
                tmp = self.M[sli, sli] * feq[0](self.stencil.get_all_velocities(i), *feq[1])
                tmp.simplify()
                eq.extend(tmp)
        return sp.Matrix(eq)

    def _conserved_moments(self, schemes):
        consm = OrderedDict()
        for i, s in enumerate(schemes):
            lo, hi = int(self.stencil.nv_ptr[i]), int(self.stencil.nv_ptr[i + 1])
            leq = list(self.EQ[lo:hi, 0])
            cm = s.get("conserved_moments", None)
            if cm is None:
                continue
            if isinstance(cm, (sp.Symbol, sp.IndexedBase)):
                cm = [cm]
            for c in cm:
                consm[c] = lo + leq.index(c)
        return consm

    def _conserved_in_front(self):
        """successive swaps (row ic <-> row c) exactly like the reference
        (pylbm/scheme.py:225-241), expressed as one permutation."""
        ns = int(self.stencil.nv_ptr[-1])
        perm = list(range(ns))
        self.permutations = []
        for ic, c in enumerate(self.consm.values()):
            self.permutations.append([ic, c])
            perm[ic], perm[c] = perm[c], perm[ic]
        self.perm = perm
        self.EQ = sp.Matrix([self.EQ[p] for p in perm])
        self.s = sp.Matrix([self.s[p] for p in perm])
        self.M = self.M.extract(perm, list(range(ns)))
        self.invM = self.invM.extract(list(range(ns)), perm)
        self.Tu = self.Tu.extract(perm, perm)
        self.Tmu = self.Tmu.extract(perm, perm)
        for ic, c in enumerate(list(self.consm.keys())):
            self.consm[c] = ic

    def __repr__(self):
        return "Scheme(dim={}, nschemes={}, nv={}, consm={}, rel_vel={})".format(
            self.dim, self.nschemes, list(self.stencil.nv), list(self.consm.keys()), self.rel_vel
        )
