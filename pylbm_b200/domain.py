"""
Discrete domain: grid coordinates, fluid/solid mask and, for every lattice
velocity, the cells whose link along that velocity crosses a boundary (with the
distance to the wall and the label of the wall).

Mirror of pylbm.domain.Domain for everything the time-step path consumes
(reference: pylbm/domain.py:252-302 constructor, 385-442 coordinates, 463-520
box faces, 523-620 elements, 622-635 clean).  It is host-side setup, but its
outputs are the input of the boundary lists, which must be bit-identical to the
reference's, so the update rules (strict `>` replacement order over the
dimensions, `alpha < distance` for solids, reset of the cells covered by an
element, final cleaning of solid cells) are followed exactly.

B200-first difference: the reference stores dense `distance`/`flag`/`normal`
arrays of shape [unvtot, nx+2v, ny+2v, nz+2v(, dim)] (float64/int64) — about
105 GB for D3Q19 at 512^3 — although only a thin layer of cells near walls ever
carries a value.  Here the per-velocity information is kept as sorted sparse
records (flat cell index in the halo-inclusive grid, distance, label); element
processing works on the element's bounding box only, one velocity at a time.
Dense `distance` / `flag` views are materialised on demand for small domains
(API compatibility and tests).  Normals are not computed (unused on the path).
"""

import copy

import numpy as np

from .geometry import Geometry
from .stencil import Stencil

__all__ = ["Domain", "SlabTopology"]


def stencil_uvel(stencil):
    """[unvtot, dim] integer array of the unique velocities (the reference's Stencil has uvx/uvy/uvz,
    stencil.py:739-767)."""
    if hasattr(stencil, "uvel"):
        return np.asarray(stencil.uvel)
    return np.asarray([stencil.uvx, stencil.uvy, stencil.uvz][: stencil.dim], dtype=np.int64).T.reshape(-1, stencil.dim)


class SlabTopology:
    """
    1-D block decomposition along x over `size` ranks, periodic in every
    direction like the reference's Cartesian communicator
    (reference: pylbm/mpi_topology.py:75-105 balanced partition,
    138-170 get_region; pylbm/domain.py:373-383 all-periodic).
    """

    def __init__(self, dim, rank=0, size=1):
        self.dim = dim
        self.rank = int(rank)
        self.size = int(size)
        self.split = np.asarray([self.size] + [1] * (dim - 1))

    @staticmethod
    def _cuts(n, nproc):
        cuts = [0]
        for i in range(nproc):
            cuts.append(cuts[-1] + n // nproc + ((n % nproc) > i))
        return cuts

    def get_region(self, nx, ny=None, nz=None):
        sizes = [n for n in (nx, ny, nz) if n is not None]
        cuts = self._cuts(int(sizes[0]), self.size)
        region = [[cuts[self.rank], cuts[self.rank + 1]]]
        for n in sizes[1:]:
            region.append([0, int(n)])
        return region

    def get_coords(self):
        return np.asarray([self.rank] + [0] * (self.dim - 1))

    @property
    def left(self):
        return (self.rank - 1) % self.size

    @property
    def right(self):
        return (self.rank + 1) % self.size


class _Records:
    """sorted sparse records (cell, dist, flag) of one unique velocity."""

    __slots__ = ("cell", "dist", "flag")

    def __init__(self, cell=None, dist=None, flag=None):
        self.cell = np.empty(0, dtype=np.int64) if cell is None else cell
        self.dist = np.empty(0, dtype=np.float64) if dist is None else dist
        self.flag = np.empty(0, dtype=np.int64) if flag is None else flag

    def keep(self, mask):
        return _Records(self.cell[mask], self.dist[mask], self.flag[mask])

    @staticmethod
    def merged(a, b):
        cell = np.concatenate([a.cell, b.cell])
        order = np.argsort(cell, kind="stable")
        return _Records(
            cell[order],
            np.concatenate([a.dist, b.dist])[order],
            np.concatenate([a.flag, b.flag])[order],
        )


class Domain:
    """
    Attributes mirrored from the reference: `dim`, `dx`, `geom`, `stencil`,
    `box_label`, `global_size`, `coords`, `coords_halo`, `shape_in`,
    `shape_halo`, `x/y/z`, `x_halo/...`, `in_or_out`, `valin`, `valout`,
    `mpi_topo`, `list_of_labels()`, and dense `distance` / `flag` on demand.
    """

    valin = 999
    valout = -1

    def __init__(self, dico, need_validation=True, topology=None, geometry_cls=None, stencil_cls=None):
        # `geometry_cls` / `stencil_cls`: the classes of an importable reference pylbm (plugin.py); the
        # builder only uses their public attributes, and elements through get_bounds / point_inside /
        # distance, so the reference's own elements (STL, cylinders, ...) work as they are
        self.geom = (geometry_cls or Geometry)(dico, need_validation=False)
        self.stencil = (stencil_cls or Stencil)(dico, need_validation=False)
        self._uvel = stencil_uvel(self.stencil)
        self.dx = dico["space_step"]
        self.dim = self.geom.dim
        self.compute_normal = False
        self.box_label = copy.copy(self.geom.box_label)

        if topology is None:
            topology = dico.get("topology", None) or SlabTopology(self.dim)
        self.mpi_topo = topology

        self.global_size = []
        self.create_coords()
        region = self.mpi_topo.get_region(*self.global_size)
        self.region = region
        for i in range(self.dim):
            if region[i][0] != 0:
                self.box_label[2 * i] = -2
            if region[i][1] != self.global_size[i]:
                self.box_label[2 * i + 1] = -2

        self.in_or_out = self.valin * np.ones(self.shape_halo)
        self._records = [_Records() for _ in range(self.stencil.unvtot)]

        self._add_box(self.box_label)
        for elem in self.geom.list_elem:
            self._add_elem(elem)
        self.clean()

    # ------------------------------------------------------------------
    # grid
    # ------------------------------------------------------------------
    def create_coords(self):
        """(reference: pylbm/domain.py:385-442)"""
        phys_box = self.geom.bounds
        for k in range(self.dim):
            npts = (phys_box[k][1] - phys_box[k][0]) / self.dx
            if not float(npts).is_integer():
                rounded = round(npts)
                if abs(rounded - npts) < self.dx:
                    npts = rounded
                else:
                    raise ValueError(
                        "The length of the box in the direction {0:d} must be a multiple of the "
                        "space step (number of points {1:.15f})".format(k, npts)
                    )
            self.global_size.append(npts)
        self.global_size = np.asarray(self.global_size, dtype="int")
        region = self.mpi_topo.get_region(*self.global_size)
        region_size = [r[1] - r[0] for r in region]

        halo_size = np.asarray(self.stencil.vmax)
        halo_beg = self.dx * (halo_size - 0.5)
        self.coords_halo = [
            np.linspace(
                phys_box[k][0] + self.dx * region[k][0] - halo_beg[k],
                phys_box[k][0] + self.dx * region[k][1] + halo_beg[k],
                region_size[k] + 2 * halo_size[k],
            )
            for k in range(self.dim)
        ]
        self.coords = [
            self.coords_halo[k][halo_size[k] : (-halo_size[k] if halo_size[k] > 0 else 1)]
            for k in range(self.dim)
        ]

    @property
    def shape_halo(self):
        return [c.size for c in self.coords_halo]

    @property
    def shape_in(self):
        return [c.size for c in self.coords]

    x = property(lambda self: self.coords[0])
    y = property(lambda self: self.coords[1])
    z = property(lambda self: self.coords[2])
    x_halo = property(lambda self: self.coords_halo[0])
    y_halo = property(lambda self: self.coords_halo[1])
    z_halo = property(lambda self: self.coords_halo[2])

    def get_bounds_halo(self):
        return (
            np.asarray([c[0] for c in self.coords_halo]),
            np.asarray([c[-1] for c in self.coords_halo]),
        )

    def get_bounds(self):
        return np.asarray([c[0] for c in self.coords]), np.asarray([c[-1] for c in self.coords])

    def list_of_labels(self):
        return np.union1d(np.unique(self.box_label), self.geom.list_of_elements_labels())

    # ------------------------------------------------------------------
    # box faces
    # ------------------------------------------------------------------
    def _add_box(self, label):
        """
        Links cut by the faces of the box (reference: pylbm/domain.py:463-520).
        For velocity k and direction d with component c = v_k[d] != 0, the
        |c| layers of interior cells next to the face the velocity points to
        get the distance (i + 1/2)/|c| and the label of the face, but only
        where this is strictly smaller than what an earlier direction set.
        """
        halo = np.asarray(self.stencil.vmax)
        shape_halo = self.shape_halo
        shape_in = self.shape_in
        self.in_or_out[:] = self.valout
        inner = tuple(slice(h, -h) if h > 0 else slice(None) for h in halo)
        self.in_or_out[inner] = self.valin

        uvel = self._uvel
        for k in range(self.stencil.unvtot):
            cand_cell, cand_dist, cand_flag = [], [], []
            for d in range(self.dim):
                c = int(uvel[k, d])
                if c < 0 and label[2 * d] != -2:
                    layers = [(halo[d] + i, -(i + 0.5) / c, label[2 * d]) for i in range(-c)]
                elif c > 0 and label[2 * d + 1] != -2:
                    layers = [
                        (halo[d] + shape_in[d] - 1 - i, (i + 0.5) / c, label[2 * d + 1])
                        for i in range(c)
                    ]
                else:
                    continue
                for index, dist, lab in layers:
                    ranges = [
                        np.arange(halo[j], halo[j] + shape_in[j]) if j != d else np.array([index])
                        for j in range(self.dim)
                    ]
                    grid = np.meshgrid(*ranges, indexing="ij")
                    cells = np.ravel_multi_index([g.ravel() for g in grid], shape_halo)
                    cand_cell.append(cells.astype(np.int64))
                    cand_dist.append(np.full(cells.size, dist))
                    cand_flag.append(np.full(cells.size, lab, dtype=np.int64))
            if not cand_cell:
                continue
            cell = np.concatenate(cand_cell)
            dist = np.concatenate(cand_dist)
            flag = np.concatenate(cand_flag)
            seq = np.arange(cell.size)
            # winner per cell: smallest distance, earliest candidate on ties
            order = np.lexsort((seq, dist, cell))
            cell, dist, flag = cell[order], dist[order], flag[order]
            first = np.ones(cell.size, dtype=bool)
            first[1:] = cell[1:] != cell[:-1]
            self._records[k] = _Records(cell[first], dist[first], flag[first])

    # ------------------------------------------------------------------
    # elements
    # ------------------------------------------------------------------
    def _add_elem(self, elem):
        """
        Add a solid (or fluid) element (reference: pylbm/domain.py:523-620),
        working on the bounding box of the element only.
        """
        vmax = np.asarray(self.stencil.vmax)
        shape_halo = self.shape_halo
        elem_bl, elem_ur = elem.get_bounds()
        phys_bl, _ = self.get_bounds_halo()
        tmp = np.array((elem_bl - phys_bl) / self.dx, int) - vmax
        nmin = np.maximum(vmax, tmp)
        tmp = np.array((elem_ur - phys_bl) / self.dx, int) + vmax + 1
        nmax = np.minimum(vmax + self.shape_in, tmp)
        if np.any(nmax <= nmin):
            return

        box = tuple(slice(lo, hi) for lo, hi in zip(nmin, nmax))
        box_shape = tuple(int(hi - lo) for lo, hi in zip(nmin, nmax))
        ioo_view = self.in_or_out[box]
        grid = np.meshgrid(
            *(self.coords_halo[d][s] for d, s in enumerate(box)), sparse=True, indexing="ij"
        )

        if not elem.isfluid:
            ind_solid = elem.point_inside(grid)
            ind_fluid = np.logical_not(ind_solid)
            ioo_view[ind_solid] = self.valout
        else:
            ind_fluid = elem.point_inside(grid)
            ind_solid = np.logical_not(ind_fluid)
            ioo_view[ind_fluid] = self.valin

        uvel = self._uvel
        for k in range(self.stencil.unvtot):
            vk = uvel[k]
            if not np.any(vk != 0):
                continue
            shifted = tuple(slice(lo + vk[d], hi + vk[d]) for d, (lo, hi) in enumerate(zip(nmin, nmax)))
            out_cells = self.in_or_out[shifted] == self.valout
            alpha, border, _ = elem.distance(grid, self.dx * vk, 1.0, False)
            indx = np.logical_and(alpha > 0, ind_fluid)
            if out_cells.size != 0:
                indx = np.logical_and(indx, out_cells)

            # local dense copy of the records of velocity k inside the box
            rec = self._records[k]
            multi = np.unravel_index(rec.cell, shape_halo)
            inside = np.ones(rec.cell.size, dtype=bool)
            for d in range(self.dim):
                inside &= (multi[d] >= nmin[d]) & (multi[d] < nmax[d])
            local = tuple(multi[d][inside] - nmin[d] for d in range(self.dim))
            dist_view = np.full(box_shape, float(self.valin))
            flag_view = np.full(box_shape, self.valin, dtype=np.int64)
            dist_view[local] = rec.dist[inside]
            flag_view[local] = rec.flag[inside]

            if elem.isfluid:
                stay_fluid = np.logical_and(np.logical_not(out_cells), ioo_view == self.valin)
                dist_view[stay_fluid] = self.valin
                flag_view[stay_fluid] = self.valin
            else:
                dist_view[ind_solid] = self.valin
                flag_view[ind_solid] = self.valin

            ind4 = np.where(indx)
            if not elem.isfluid:
                ind3 = np.where(alpha[ind4] < dist_view[ind4])[0]
            else:
                ind3 = np.where(
                    np.logical_or(alpha[ind4] > dist_view[ind4], dist_view[ind4] == self.valin)
                )[0]
            ind = tuple(i[ind3] for i in ind4)
            dist_view[ind] = alpha[ind]
            flag_view[ind] = border[ind]

            # back to sparse records
            touched = np.where(np.logical_or(dist_view != self.valin, flag_view != self.valin))
            cells = np.ravel_multi_index(
                tuple(t + nmin[d] for d, t in enumerate(touched)), shape_halo
            ).astype(np.int64)
            new = _Records(cells, dist_view[touched], flag_view[touched])
            self._records[k] = _Records.merged(rec.keep(~inside), new)

    def clean(self):
        """drop what was computed on cells that ended up solid
        (reference: pylbm/domain.py:622-635)."""
        ioo = self.in_or_out.ravel()
        for k, rec in enumerate(self._records):
            bad = np.logical_and(rec.dist > 0, ioo[rec.cell] == self.valout)
            if bad.any():
                self._records[k] = rec.keep(~bad)

    # ------------------------------------------------------------------
    # queries used by the boundary-list builder
    # ------------------------------------------------------------------
    def cells_with_flag(self, k, label):
        """
        multi-indices (tuple of arrays, C order of the halo grid) and distances
        of the cells whose link along unique velocity k is cut by a wall
        labelled `label`  ==  np.where(flag[k] == label) of the reference.
        """
        rec = self._records[k]
        sel = rec.flag == label
        cells = rec.cell[sel]
        return np.unravel_index(cells, self.shape_halo), rec.dist[sel]

    # ------------------------------------------------------------------
    # dense views (small domains / tests)
    # ------------------------------------------------------------------
    def _dense(self, field, dtype):
        out = np.full([self.stencil.unvtot] + self.shape_halo, self.valin, dtype=dtype)
        for k, rec in enumerate(self._records):
            out[k].ravel()[rec.cell] = getattr(rec, field)
        return out

    @property
    def distance(self):
        return self._dense("dist", np.float64)

    @property
    def flag(self):
        return self._dense("flag", np.int64)

    @property
    def normal(self):
        raise NotImplementedError("normal vectors are not computed (unused on the time-step path)")

    def __repr__(self):
        return "Domain(dim={}, dx={}, shape_in={}, region={}, labels={})".format(
            self.dim, self.dx, self.shape_in, self.region, self.box_label
        )
