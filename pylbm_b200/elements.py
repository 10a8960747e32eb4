"""
Geometrical elements that can be added to / removed from the box.

Host-side setup code (not on the per-step path) but its outputs feed the
boundary lists, which are part of the parity contract: the distances returned
here become the Bouzidi coefficients.  The floating-point evaluation order of
the ray/quadric and ray/segment intersections therefore follows the reference
formulas (reference: pylbm/elements/utils.py:21-277, circle.py:92-151,
ellipse.py, parallelogram.py, triangle.py, sphere.py:94-155, ellipsoid.py) so
that the values are bit-identical.  Normal vectors are not produced: nothing on
the time-step path reads them (reference: pylbm/boundary.py:208 stores them and
never uses them).

Supported: Circle, Ellipse, Parallelogram, Triangle (2D); Sphere, Ellipsoid (3D).
"""

import numpy as np

__all__ = ["Element", "Circle", "Ellipse", "Parallelogram", "Triangle", "Sphere", "Ellipsoid"]

_HUGE = 1.0e16


class Element:
    """
    Base class: `label` (one per edge), `isfluid`, `dim`, and the three
    methods used by the domain builder: get_bounds, point_inside, distance.
    """

    number_of_bounds = -1
    dim = 0

    def __init__(self, label, isfluid):
        self.isfluid = isfluid
        self.label = [label] * self.number_of_bounds if isinstance(label, int) else label

    def test_label(self):
        return len(self.label) == self.number_of_bounds

    def get_bounds(self):
        raise NotImplementedError

    def point_inside(self, grid):
        raise NotImplementedError

    def distance(self, grid, v, dmax=None, normal=False):
        """returns (alpha, border, None): distance along v (in units of |v|)
        to the edge, and the label of that edge; -1 where there is none."""
        raise NotImplementedError

    def __repr__(self):
        kind = "fluid" if self.isfluid else "solid"
        return "{}(label={}, {})".format(self.__class__.__name__, self.label, kind)


# --------------------------------------------------------------------------
# ray / conic intersections
# --------------------------------------------------------------------------
def _select(d, dmax, label, shape):
    alpha = -np.ones(shape)
    border = -np.ones(shape)
    keep = d > 0 if dmax is None else np.logical_and(d > 0, d <= dmax)
    alpha[keep] = d[keep]
    border[keep] = label
    return alpha, border


def _ray_conic_2d(x, y, v, center, v1, v2, dmax, label):
    """ellipse spanned by v1, v2 around center; ray (x, y) + d v."""
    X = x - center[0]
    Y = y - center[1]
    vx2 = v1[0] ** 2 + v2[0] ** 2
    vy2 = v1[1] ** 2 + v2[1] ** 2
    vxy = 2 * (v1[0] * v1[1] + v2[0] * v2[1])
    rhs = (v1[0] * v2[1] - v1[1] * v2[0]) ** 2
    a = v[0] ** 2 * vy2 + v[1] ** 2 * vx2 - v[0] * v[1] * vxy
    b = 2 * X * v[0] * vy2 + 2 * Y * v[1] * vx2 - (X * v[1] + Y * v[0]) * vxy
    c = X**2 * vy2 + Y**2 * vx2 - X * Y * vxy - rhs
    delta = b**2 - 4 * a * c
    ok = delta >= 0
    delta[ok] = np.sqrt(delta[ok])
    shape = ok.shape
    d1 = _HUGE * np.ones(shape)
    d2 = _HUGE * np.ones(shape)
    if a != 0:
        d1[ok] = (-b[ok] - delta[ok]) / (2 * a)
        d2[ok] = (-b[ok] + delta[ok]) / (2 * a)
    d1[d1 < 0] = _HUGE
    d2[d2 < 0] = _HUGE
    d = -np.ones(shape)
    d[ok] = np.minimum(d1[ok], d2[ok])
    d[d == _HUGE] = -1
    return _select(d, dmax, label, shape)


def _quadric_coefficients(v1, v2, v3):
    v12 = np.cross(v1, v2)
    v23 = np.cross(v2, v3)
    v31 = np.cross(v3, v1)
    rhs = np.inner(v1, v23) ** 2
    cxx = v12[0] ** 2 + v23[0] ** 2 + v31[0] ** 2
    cyy = v12[1] ** 2 + v23[1] ** 2 + v31[1] ** 2
    czz = v12[2] ** 2 + v23[2] ** 2 + v31[2] ** 2
    cxy = 2 * (v12[0] * v12[1] + v23[0] * v23[1] + v31[0] * v31[1])
    cyz = 2 * (v12[1] * v12[2] + v23[1] * v23[2] + v31[1] * v31[2])
    czx = 2 * (v12[2] * v12[0] + v23[2] * v23[0] + v31[2] * v31[0])
    return cxx, cyy, czz, cxy, cyz, czx, rhs


def _ray_quadric_3d(x, y, z, v, center, v1, v2, v3, dmax, label):
    """ellipsoid spanned by v1, v2, v3 around center; ray (x, y, z) + d v."""
    shape = (x.size, y.size, z.size)
    X = x - center[0]
    Y = y - center[1]
    Z = z - center[2]
    cxx, cyy, czz, cxy, cyz, czx, rhs = _quadric_coefficients(v1, v2, v3)
    a = (
        cxx * v[0] ** 2
        + cyy * v[1] ** 2
        + czz * v[2] ** 2
        + cxy * v[0] * v[1]
        + cyz * v[1] * v[2]
        + czx * v[2] * v[0]
    )
    b = (
        (2 * cxx * v[0] + cxy * v[1] + czx * v[2]) * X
        + (2 * cyy * v[1] + cyz * v[2] + cxy * v[0]) * Y
        + (2 * czz * v[2] + czx * v[0] + cyz * v[1]) * Z
    )
    c = cxx * X**2 + cyy * Y**2 + czz * Z**2 + cxy * X * Y + cyz * Y * Z + czx * Z * X - rhs
    delta = b**2 - 4 * a * c
    ok = delta >= 0
    delta[ok] = np.sqrt(delta[ok])
    d1 = _HUGE * np.ones(shape)
    d2 = _HUGE * np.ones(shape)
    d1[ok] = (-b[ok] - delta[ok]) / (2 * a)
    d2[ok] = (-b[ok] + delta[ok]) / (2 * a)
    d1[d1 < 0] = _HUGE
    d2[d2 < 0] = _HUGE
    d = -np.ones(shape)
    d[ok] = np.minimum(d1[ok], d2[ok])
    d[d == _HUGE] = -1
    return _select(d, dmax, label, shape)


def _ray_segments_2d(x, y, v, origins, edges, dmax, labels):
    """
    closest intersection of the rays (x, y) + d v with the segments
    origins[i] + s edges[i], s in [0, 1].
    """
    shape = (x.size, y.size) if (x.shape[1] == 1 and y.shape[0] == 1) else x.shape
    alpha = _HUGE * np.ones(shape)
    border = -np.ones(shape)
    for p, e, lab in zip(origins, edges, labels):
        det = v[1] * e[0] - v[0] * e[1]
        if det == 0:
            continue  # ray parallel to this edge
        invdet = 1.0 / det
        c1 = p[0] - x
        c2 = p[1] - y
        along_ray = (-e[1] * c1 + e[0] * c2) * invdet
        along_edge = (-v[1] * c1 + v[0] * c2) * invdet
        on_edge = np.logical_and(along_edge >= 0, along_edge <= 1)
        if dmax is None:
            hit = np.logical_and(along_ray > 0, on_edge)
        else:
            hit = np.logical_and(np.logical_and(along_ray > 0, along_ray <= dmax), on_edge)
        closer = np.where(np.logical_and(alpha > along_ray, hit))
        alpha[closer] = along_ray[closer]
        border[closer] = lab
    alpha[alpha == _HUGE] = -1.0
    return alpha, border


# --------------------------------------------------------------------------
# 2D elements
# --------------------------------------------------------------------------
class Circle(Element):
    """Circle(center, radius, label=0, isfluid=False) (reference: elements/circle.py)"""

    number_of_bounds = 1
    dim = 2

    def __init__(self, center, radius, label=0, isfluid=False):
        self.center = np.asarray(center)
        if radius < 0:
            raise ValueError("The radius of the circle should be positive")
        self.radius = radius
        super().__init__(label, isfluid)

    def get_bounds(self):
        return self.center - self.radius, self.center + self.radius

    def point_inside(self, grid):
        x, y = grid
        rel = [x - self.center[0], y - self.center[1]]
        return (rel[0] ** 2 + rel[1] ** 2) <= self.radius**2

    def distance(self, grid, v, dmax=None, normal=False):
        x, y = grid
        v1 = self.radius * np.array([1, 0])
        v2 = self.radius * np.array([0, 1])
        alpha, border = _ray_conic_2d(x, y, v, self.center, v1, v2, dmax, self.label[0])
        return alpha, border, None


class Ellipse(Element):
    """Ellipse(center, v1, v2, label=0, isfluid=False) (reference: elements/ellipse.py)"""

    number_of_bounds = 1
    dim = 2

    def __init__(self, center, v1, v2, label=0, isfluid=False):
        self.center = np.asarray(center)
        if abs(v1[0] * v2[0] + v1[1] * v2[1]) > 1.0e-14:
            raise ValueError("The vectors of the ellipse are not orthogonal")
        self.v1 = np.asarray(v1)
        self.v2 = np.asarray(v2)
        super().__init__(label, isfluid)

    def get_bounds(self):
        r = max(np.linalg.norm(self.v1), np.linalg.norm(self.v2))
        return self.center - r, self.center + r

    def point_inside(self, grid):
        x, y = grid
        X = x - self.center[0]
        Y = y - self.center[1]
        vx2 = self.v1[0] ** 2 + self.v2[0] ** 2
        vy2 = self.v1[1] ** 2 + self.v2[1] ** 2
        vxy = 2 * (self.v1[0] * self.v1[1] + self.v2[0] * self.v2[1])
        det = self.v1[0] * self.v2[1] - self.v1[1] * self.v2[0]
        return X**2 * vy2 + Y**2 * vx2 - X * Y * vxy <= det**2

    def distance(self, grid, v, dmax=None, normal=False):
        x, y = grid
        alpha, border = _ray_conic_2d(x, y, v, self.center, self.v1, self.v2, dmax, self.label[0])
        return alpha, border, None


class _TwoVectors(Element):
    """shapes defined by a corner and two edge vectors."""

    dim = 2

    def __init__(self, point, vecta, vectb, label=0, isfluid=False):
        self.point = np.asarray(point)
        self.v1 = np.asarray(vecta)
        self.v2 = np.asarray(vectb)
        super().__init__(label, isfluid)

    def get_bounds(self):
        corners = np.asarray(
            [self.point, self.point + self.v1, self.point + self.v1 + self.v2, self.point + self.v2]
        )
        return np.min(corners, axis=0), np.max(corners, axis=0)

    def _barycentric(self, grid):
        x, y = grid
        rel = [x - self.point[0], y - self.point[1]]
        invdelta = 1.0 / (self.v1[0] * self.v2[1] - self.v1[1] * self.v2[0])
        u = (rel[0] * self.v2[1] - rel[1] * self.v2[0]) * invdelta
        w = (rel[1] * self.v1[0] - rel[0] * self.v1[1]) * invdelta
        return u, w

    def _segments(self):
        raise NotImplementedError

    def distance(self, grid, v, dmax=None, normal=False):
        x, y = grid
        origins, edges = self._segments()
        alpha, border = _ray_segments_2d(
            x - self.point[0], y - self.point[1], v, origins, edges, dmax, self.label
        )
        return alpha, border, None


class Parallelogram(_TwoVectors):
    """Parallelogram(point, vecta, vectb, label=0, isfluid=False)
    (reference: elements/parallelogram.py); 4 labelled edges."""

    number_of_bounds = 4

    def point_inside(self, grid):
        u, w = self._barycentric(grid)
        return np.logical_and(np.logical_and(u >= 0, w >= 0), np.logical_and(u <= 1, w <= 1))

    def _segments(self):
        return [[0, 0], [0, 0], self.v2, self.v1], [self.v1, self.v2, self.v1, self.v2]


class Triangle(_TwoVectors):
    """Triangle(point, vecta, vectb, label=0, isfluid=False)
    (reference: elements/triangle.py); 3 labelled edges."""

    number_of_bounds = 3

    def point_inside(self, grid):
        u, w = self._barycentric(grid)
        return np.logical_and(np.logical_and(u >= 0, w >= 0), u + w <= 1)

    def _segments(self):
        return [[0, 0], [0, 0], self.v1], [self.v1, self.v2, self.v2 - self.v1]


# --------------------------------------------------------------------------
# 3D elements
# --------------------------------------------------------------------------
class Sphere(Element):
    """Sphere(center, radius, label=0, isfluid=False) (reference: elements/sphere.py)"""

    number_of_bounds = 1
    dim = 3

    def __init__(self, center, radius, label=0, isfluid=False):
        self.center = np.asarray(center)
        if radius < 0:
            raise ValueError("The radius of the sphere should be positive")
        self.radius = radius
        super().__init__(label, isfluid)

    def get_bounds(self):
        return self.center - self.radius, self.center + self.radius

    def point_inside(self, grid):
        x, y, z = grid
        rel = [x - self.center[0], y - self.center[1], z - self.center[2]]
        return (rel[0] ** 2 + rel[1] ** 2 + rel[2] ** 2) <= self.radius**2

    def distance(self, grid, v, dmax=None, normal=False):
        x, y, z = grid
        v1 = self.radius * np.array([1, 0, 0])
        v2 = self.radius * np.array([0, 1, 0])
        v3 = self.radius * np.array([0, 0, 1])
        alpha, border = _ray_quadric_3d(x, y, z, v, self.center, v1, v2, v3, dmax, self.label[0])
        return alpha, border, None


class Ellipsoid(Element):
    """Ellipsoid(center, v1, v2, v3, label=0, isfluid=False)
    (reference: elements/ellipsoid.py)"""

    number_of_bounds = 1
    dim = 3

    def __init__(self, center, v1, v2, v3, label=0, isfluid=False):
        self.center = np.asarray(center)
        dots = [abs(sum(a[i] * b[i] for i in range(3))) for a, b in ((v1, v2), (v2, v3), (v3, v1))]
        if max(dots) > 1.0e-14:
            raise ValueError("The vectors of the ellipsoid are not orthogonal")
        self.v1, self.v2, self.v3 = np.asarray(v1), np.asarray(v2), np.asarray(v3)
        super().__init__(label, isfluid)

    def get_bounds(self):
        r = max(np.linalg.norm(self.v1), np.linalg.norm(self.v2), np.linalg.norm(self.v3))
        return self.center - r, self.center + r

    def point_inside(self, grid):
        x, y, z = grid
        X = x - self.center[0]
        Y = y - self.center[1]
        Z = z - self.center[2]
        cxx, cyy, czz, cxy, cyz, czx, rhs = _quadric_coefficients(self.v1, self.v2, self.v3)
        return (
            cxx * X**2 + cyy * Y**2 + czz * Z**2 + cxy * X * Y + cyz * Y * Z + czx * Z * X
        ) <= rhs

    def distance(self, grid, v, dmax=None, normal=False):
        x, y, z = grid
        alpha, border = _ray_quadric_3d(
            x, y, z, v, self.center, self.v1, self.v2, self.v3, dmax, self.label[0]
        )
        return alpha, border, None
