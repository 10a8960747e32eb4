"""
Lattice velocities and stencils.

Host-side mirror of the reference's velocity numbering convention
(reference: pylbm/stencil.py:285-373 for num <-> coordinates,
pylbm/stencil.py:643-696 for the per-scheme bookkeeping).  The numbering is
part of the parity contract: the population index k of a boundary-list entry
and the row order of M follow it.

The implementation is table driven: the velocities of every "shell" are
enumerated once in numbering order and looked up, instead of the reference's
closed-form / brute-force search.
"""

import itertools
from functools import lru_cache

import numpy as np

__all__ = ["Velocity", "Stencil", "OneStencil"]


def _shell_2d(n):
    """velocities with max(|vx|,|vy|) == n in numbering order."""
    if n == 0:
        return [(0, 0)]
    out = [(n, 0), (0, n), (-n, 0), (0, -n), (n, n), (-n, n), (-n, -n), (n, -n)]
    for j in range(1, n):
        out += [
            (n, j), (j, n), (-j, n), (-n, j),
            (-n, -j), (-j, -n), (j, -n), (n, -j),
        ]
    return out


def _signed(a):
    return (a,) if a == 0 else (a, -a)


def _shell_3d(k):
    """velocities with max(|v|) == k in numbering order (3D)."""
    out = []
    for i in range(k + 1):
        for j in range(i + 1):
            for kk, ii, jj in sorted(set(itertools.permutations((k, i, j)))):
                for a in _signed(kk):
                    for b in _signed(ii):
                        for c in _signed(jj):
                            out.append((a, b, c))
    return out


@lru_cache(maxsize=None)
def _table(dim, nshell):
    """(list of velocities in numbering order up to shell nshell, reverse map)"""
    vel = []
    for n in range(nshell + 1):
        if dim == 1:
            vel += [(0,)] if n == 0 else [(n,), (-n,)]
        elif dim == 2:
            vel += _shell_2d(n)
        else:
            vel += _shell_3d(n)
    return vel, {v: i for i, v in enumerate(vel)}


def _num_to_coord(dim, num):
    nshell = 1
    while True:
        vel, _ = _table(dim, nshell)
        if num < len(vel):
            return vel[num]
        nshell *= 2


def _coord_to_num(dim, coord):
    nshell = max(1, max(abs(c) for c in coord))
    return _table(dim, nshell)[1][tuple(int(c) for c in coord)]


class Velocity:
    """
    One lattice velocity, defined either by (dim, num) or by its components.

    Mirrors pylbm.stencil.Velocity (reference: pylbm/stencil.py:119-373):
    attributes dim, num, vx, vy, vz, v, v_full and get_symmetric(axis).
    """

    def __init__(self, dim=None, num=None, vx=None, vy=None, vz=None):
        if dim is None:
            if vz is not None:
                dim = 3
            elif vy is not None:
                dim = 2
            elif vx is not None:
                dim = 1
            else:
                raise ValueError("a velocity needs (dim, num) or its components")
        self.dim = dim
        if vx is None:
            if num is None:
                raise ValueError("a velocity needs (dim, num) or its components")
            coord = _num_to_coord(dim, int(num))
        else:
            coord = tuple(int(c) for c in (vx, vy, vz)[:dim])
        if num is None:
            num = _coord_to_num(dim, coord)
        self.num = int(num)
        full = list(coord) + [None] * (3 - dim)
        self.vx, self.vy, self.vz = full

    @property
    def v(self):
        return [self.vx, self.vy, self.vz][: self.dim]

    @property
    def v_full(self):
        return [c if c is not None else 0 for c in (self.vx, self.vy, self.vz)]

    def __repr__(self):
        return "({}: {})".format(self.num, ", ".join(str(c) for c in self.v))

    __str__ = __repr__

    def get_symmetric(self, axis=None):
        """
        Symmetric velocity: through the origin (axis None) or keeping the
        component along `axis` (reference: pylbm/stencil.py:241-282).
        """
        if axis is not None and not 0 <= axis < self.dim:
            raise ValueError("axis must be lower than the dimension of the velocity")
        comp = [-c for c in self.v]
        if axis is not None:
            comp[axis] = self.v[axis]
        comp += [None] * (3 - self.dim)
        return Velocity(vx=comp[0], vy=comp[1], vz=comp[2])


class OneStencil:
    """velocities of one elementary scheme (reference: pylbm/stencil.py:376-441)"""

    def __init__(self, v, nv):
        self.v = v
        self.nv = nv
        self.num = np.array([vk.num for vk in v], dtype=int)
        self.vx = np.array([vk.vx for vk in v])
        self.vy = np.array([vk.vy for vk in v])
        self.vz = np.array([vk.vz for vk in v])


class _PerScheme:
    """tiny helper giving `stencil.vx[k]`-style item access."""

    def __init__(self, getter):
        self._getter = getter

    def __getitem__(self, k):
        return self._getter(k)


class Stencil(list):
    """
    The velocities of all the elementary schemes of a dictionary.

    Mirror of pylbm.stencil.Stencil (reference: pylbm/stencil.py:444-835):
    `dim`, `nstencils`, `unique_velocities` (sorted by number), `v[k]`,
    `nv[k]`, `nv_ptr`, `num[k]`, `unum`, `unum2index`, `uvx/uvy/uvz`, `vmax`,
    `get_all_velocities()`, `get_symmetric()`.
    """

    def __init__(self, dico, need_validation=True):
        super().__init__()
        self.dim = self.extract_dim(dico)
        schemes_velocities = [np.asarray(s["velocities"]) for s in dico["schemes"]]
        self.nstencils = len(schemes_velocities)

        unique = np.empty(0, dtype=np.int32)
        for vel in schemes_velocities:
            unique = np.union1d(unique, vel)
        self.unique_velocities = np.asarray(
            [Velocity(dim=self.dim, num=int(i)) for i in unique], dtype=object
        )
        self._unum = np.asarray([int(i) for i in unique], dtype=int)

        self.v, self.nv, ptr = [], [], [0]
        for vel in schemes_velocities:
            pos = np.searchsorted(self._unum, vel)
            self.v.append(self.unique_velocities[pos])
            self.nv.append(len(vel))
            ptr.append(ptr[-1] + len(vel))
        self.nv_ptr = np.asarray(ptr)

        self.num2index = []
        for k in range(self.nstencils):
            self.num2index.extend(int(vk.num) for vk in self.v[k])

        self.unum2index = -1000 + np.zeros(np.max(self._unum) + 1, dtype=np.int32)
        self.unum2index[self._unum] = np.arange(self._unum.size)

        for k in range(self.nstencils):
            self.append(OneStencil(self.v[k], self.nv[k]))

        self.num = _PerScheme(lambda k: self[k].num)
        self.vx = _PerScheme(lambda k: self[k].vx)
        self.vy = _PerScheme(lambda k: self[k].vy)
        self.vz = _PerScheme(lambda k: self[k].vz)

    @staticmethod
    def extract_dim(dico):
        dim = dico.get("dim", None)
        if not dim:
            box = dico["box"]
            dim = 1 + ("y" in box and box["y"] is not None)
            if dim == 2:
                dim += "z" in box and box["z"] is not None
        return dim

    @property
    def unvtot(self):
        return self.unique_velocities.size

    @property
    def unum(self):
        return self._unum.copy()

    def _ucomp(self, d):
        return np.array([vk.v_full[d] if d < self.dim else None for vk in self.unique_velocities])

    @property
    def uvx(self):
        return self._ucomp(0)

    @property
    def uvy(self):
        return self._ucomp(1)

    @property
    def uvz(self):
        return self._ucomp(2)

    @property
    def uvel(self):
        """integer array (unvtot, dim) of the unique velocities."""
        return np.array([vk.v for vk in self.unique_velocities], dtype=int).reshape(-1, self.dim)

    @property
    def vmax(self):
        return np.max(self.uvel, axis=0)

    @property
    def vmin(self):
        return np.min(self.uvel, axis=0)

    @property
    def vmax_full(self):
        out = np.zeros(3, dtype=int)
        out[: self.dim] = self.vmax
        return out

    def get_all_velocities(self, scheme_id=None):
        """(nv, dim) integer array (reference: pylbm/stencil.py:783-815)."""
        ids = range(self.nstencils) if scheme_id is None else [scheme_id]
        rows = [vk.v for k in ids for vk in self.v[k]]
        return np.asarray(rows, dtype=int).reshape(-1, self.dim)

    def get_symmetric(self, axis=None):
        """
        index (in the global population numbering) of the symmetric velocity
        inside the same elementary scheme (reference: pylbm/stencil.py:817-835).
        """
        ksym = np.empty(self.nv_ptr[-1], dtype=np.int32)
        k = 0
        for n, v in enumerate(self.v):
            local = self.num2index[self.nv_ptr[n] : self.nv_ptr[n + 1]]
            for vk in v:
                ksym[k] = local.index(vk.get_symmetric(axis).num) + self.nv_ptr[n]
                k += 1
        return ksym

    def is_symmetric(self):
        for n, v in enumerate(self.v):
            local = self.num2index[self.nv_ptr[n] : self.nv_ptr[n + 1]]
            for vk in v:
                if vk.get_symmetric().num not in local:
                    return False
        return True

    def __repr__(self):
        lines = ["Stencil: dim {} / {} scheme(s)".format(self.dim, self.nstencils)]
        for k in range(self.nstencils):
            lines.append("  scheme {}: {}".format(k, list(self.v[k])))
        return "\n".join(lines)
