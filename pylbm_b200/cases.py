"""
Scheme dictionaries of the benchmark / parity workloads (BASELINE.json configs).

Each builder returns a plain pylbm dictionary.  `mod` is the namespace providing
the boundary-method classes and the geometric elements (`mod.bc.BounceBack`,
`mod.Circle`, ...): the package itself by default, or the reference `pylbm`
module when tools/make_golden.py generates fixtures with the very same
dictionary.  `perturb=seed` replaces the uniform initial state by the seeded
random perturbation used by the parity tests so that every moment
is exercised.

C1 / C2 follow the reference demos demo/2D/lid_driven_cavity.py:57-232 and
demo/2D/Karman_vortex_street.py:70-227 (Geier central moments, relative
velocity); C3 follows demo/2D/shallow_water.py:62-122; C4 (D3Q19 MRT) and C5
(D3Q27 tensor-product central moments) have no demo in the reference and are
authored here (d'Humieres et al. 2002 moment set for D3Q19).
"""

import numpy as np
import sympy as sp

X, Y, Z = sp.symbols("X, Y, Z")
RHO, QX, QY, QZ = sp.symbols("rho, qx, qy, qz")
LA = sp.symbols("lambda", constants=True)


def _default_mod():
    import pylbm_b200

    return pylbm_b200


class _Wave(int):
    """perturb='wave': smooth perturbation that depends on the COORDINATES only (not on the local
    array shape), so that every slab of a decomposed run initialises the same global field."""

    def __add__(self, other):
        return _Wave(int(self) + other)


WAVE = _Wave(0)


def _perturbed(seed, base, amp=0.01):
    """callable init: base + amp*U(-1,1), one independent stream per moment (seeded NumPy
    generator over the local array), or a coordinate-based wave for seed = WAVE + i."""

    def init(*coords):
        shape = np.broadcast(*coords).shape
        if isinstance(seed, _Wave):
            field = np.ones(shape)
            for i, c in enumerate(coords):
                field = field * np.cos(2 * np.pi * (2 + i + int(seed)) * c + 0.7 * (i + 1) + int(seed))
            return base + amp * field
        rng = np.random.default_rng(seed)
        return base + amp * rng.uniform(-1.0, 1.0, size=shape)

    return init


# --------------------------------------------------------------------------
# D2Q9
# --------------------------------------------------------------------------
def _d2q9_geier(space_step, mu, zeta, la=1.0, rho_o=1.0):
    ux, uy = QX / RHO, QY / RHO
    polynomials = [1, X, Y, X**2 + Y**2, X * Y**2, Y * X**2, X**2 * Y**2, X**2 - Y**2, X * Y]
    equilibrium = [
        RHO,
        QX,
        QY,
        RHO * (ux**2 + uy**2) + 2 / 3 * RHO * LA**2,
        QX * (LA**2 / 3 + uy**2),
        QY * (LA**2 / 3 + ux**2),
        RHO * (LA**2 / 3 + ux**2) * (LA**2 / 3 + uy**2),
        RHO * (ux**2 - uy**2),
        RHO * ux * uy,
    ]
    dummy = 3.0 / (la * rho_o * space_step)
    s_1 = 1 / (0.5 + dummy * (zeta - 2 * mu / 3))
    s_2 = 1 / (0.5 + dummy * mu)
    return polynomials, equilibrium, [0.0, 0.0, 0.0, s_1, s_1, s_1, s_1, s_2, s_2]


def _wall_value_2d(f, m, x, y, rho_o, ux_o):
    m[RHO] = rho_o
    m[QX] = rho_o * ux_o
    m[QY] = 0.0


def lid_cavity_d2q9(n=256, mod=None, perturb=None, generator="cuda"):
    """C1: D2Q9 lid-driven cavity n x n, Geier moments, relative velocity, Bouzidi bounce-back."""
    mod = mod or _default_mod()
    dx = 1.0 / n
    la, rho_o, lid, mu = 1.0, 1.0, 0.05, 5.0e-6
    pol, eq, s = _d2q9_geier(dx, mu, 100 * mu, la, rho_o)
    init = {RHO: rho_o, QX: 0.0, QY: 0.0}
    if perturb is not None:
        init = {RHO: _perturbed(perturb, rho_o), QX: _perturbed(perturb + 1, 0.0), QY: _perturbed(perturb + 2, 0.0)}
    return {
        "box": {"x": [0.0, 1.0], "y": [0.0, 1.0], "label": [0, 0, 0, 1]},
        "space_step": dx,
        "scheme_velocity": la,
        "schemes": [
            {
                "velocities": list(range(9)),
                "polynomials": pol,
                "relaxation_parameters": s,
                "equilibrium": eq,
                "conserved_moments": [RHO, QX, QY],
            }
        ],
        "parameters": {LA: la},
        "init": init,
        "boundary_conditions": {
            0: {"method": {0: mod.bc.BouzidiBounceBack}},
            1: {"method": {0: mod.bc.BouzidiBounceBack}, "value": (_wall_value_2d, (rho_o, lid))},
        },
        "relative_velocity": [QX / RHO, QY / RHO],
        "generator": generator,
    }


def karman_d2q9(nx=4096, ny=1024, mod=None, perturb=None, generator="cuda", relative_velocity=True,
                radius=1.0 / 16, cx=None):
    """C2: D2Q9 Karman vortex street nx x ny behind a circular obstacle (Bouzidi bounce-back
    on inlet, walls and obstacle, Neumann outlet)."""
    mod = mod or _default_mod()
    dx = 1.0 / ny
    length = nx * dx
    la, rho_o, u_o, mu = 1.0, 1.0, 0.05, 5.0e-6
    pol, eq, s = _d2q9_geier(dx, mu, 10 * mu, la, rho_o)
    init = {RHO: rho_o, QX: rho_o * u_o, QY: 0.0}
    if perturb is not None:
        init = {
            RHO: _perturbed(perturb, rho_o),
            QX: _perturbed(perturb + 1, rho_o * u_o),
            QY: _perturbed(perturb + 2, 0.0),
        }
    dico = {
        "box": {"x": [0.0, length], "y": [0.0, 1.0], "label": [0, 1, 0, 0]},
        "elements": [mod.Circle([0.15 * length if cx is None else cx, 0.5 + 2 * dx], radius, label=2)],
        "space_step": dx,
        "scheme_velocity": la,
        "schemes": [
            {
                "velocities": list(range(9)),
                "polynomials": pol,
                "relaxation_parameters": s,
                "equilibrium": eq,
                "conserved_moments": [RHO, QX, QY],
            }
        ],
        "parameters": {LA: la},
        "init": init,
        "boundary_conditions": {
            0: {"method": {0: mod.bc.BouzidiBounceBack}, "value": (_wall_value_2d, (rho_o, u_o))},
            1: {"method": {0: mod.bc.NeumannX}},
            2: {"method": {0: mod.bc.BouzidiBounceBack}},
        },
        "generator": generator,
    }
    if relative_velocity:
        dico["relative_velocity"] = [QX / RHO, QY / RHO]
    return dico


# --------------------------------------------------------------------------
# D2Q4 x 3 vectorial shallow water
# --------------------------------------------------------------------------
H = sp.symbols("h")
G = sp.symbols("g", constants=True)
_SIG = sp.symbols("sigma_0, sigma_1, sigma_2, sigma_3", constants=True)


def _h_bump(x, y, xmin, xmax, ymin, ymax):
    cx, cy = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
    return 1 + 0.5 * ((x - cx) ** 2 + (y - cy) ** 2 < 0.1**2)


def shallow_water_d2q4(n=4096, mod=None, perturb=None, generator="cuda"):
    """C3: three coupled D2Q4 schemes (h, qx, qy), relative velocity, fully periodic box."""
    xmin, xmax, ymin, ymax = -1.0, 1.0, -1.0, 1.0
    dx = (xmax - xmin) / n
    s_h = [0.0, 1 / (0.5 + _SIG[0]), 1 / (0.5 + _SIG[0]), 1 / (0.5 + _SIG[1])]
    s_q = [0.0, 1 / (0.5 + _SIG[2]), 1 / (0.5 + _SIG[2]), 1 / (0.5 + _SIG[3])]
    vel = list(range(1, 5))
    pol = [1, X, Y, X**2 - Y**2]
    init = {H: (_h_bump, (xmin, xmax, ymin, ymax)), QX: 0.0, QY: 0.0}
    if perturb is not None:
        init = {H: _perturbed(perturb, 1.0), QX: _perturbed(perturb + 1, 0.0), QY: _perturbed(perturb + 2, 0.0)}
    return {
        "parameters": {LA: 4, G: 1.0, _SIG[0]: 1.0e-3, _SIG[1]: 0.5, _SIG[2]: 1.0e-1, _SIG[3]: 0.5},
        "box": {"x": [xmin, xmax], "y": [ymin, ymax], "label": -1},
        "space_step": dx,
        "scheme_velocity": LA,
        "schemes": [
            {"velocities": vel, "conserved_moments": H, "polynomials": pol,
             "relaxation_parameters": s_h, "equilibrium": [H, QX, QY, 0.0]},
            {"velocities": vel, "conserved_moments": QX, "polynomials": pol,
             "relaxation_parameters": s_q, "equilibrium": [QX, QX**2 / H + G * H**2 / 2, QX * QY / H, 0.0]},
            {"velocities": vel, "conserved_moments": QY, "polynomials": pol,
             "relaxation_parameters": s_q, "equilibrium": [QY, QX * QY / H, QY**2 / H + G * H**2 / 2, 0.0]},
        ],
        "init": init,
        "relative_velocity": [QX / H, QY / H],
        "generator": generator,
    }


# --------------------------------------------------------------------------
# D3Q19 MRT lid-driven cavity
# --------------------------------------------------------------------------
def _wall_value_3d(f, m, x, y, z, rho_o, ux_o):
    m[RHO] = rho_o
    m[QX] = rho_o * ux_o
    m[QY] = 0.0
    m[QZ] = 0.0


def lid_cavity_d3q19(n=512, mod=None, perturb=None, generator="cuda", nu=0.02):
    """C4: D3Q19 MRT (d'Humieres moments) lid-driven cavity n^3, bounce-back walls, moving lid z+."""
    mod = mod or _default_mod()
    dx = 1.0 / n
    rho_o, lid = 1.0, 0.05
    r = X**2 + Y**2 + Z**2
    polynomials = [
        1, 19 * r - 30, (21 * r**2 - 53 * r + 24) / 2,
        X, (5 * r - 9) * X, Y, (5 * r - 9) * Y, Z, (5 * r - 9) * Z,
        3 * X**2 - r, (3 * r - 5) * (3 * X**2 - r),
        Y**2 - Z**2, (3 * r - 5) * (Y**2 - Z**2),
        X * Y, Y * Z, Z * X,
        (Y**2 - Z**2) * X, (Z**2 - X**2) * Y, (X**2 - Y**2) * Z,
    ]
    j2 = QX**2 + QY**2 + QZ**2
    pxx = 2 * QX**2 - QY**2 - QZ**2
    pww = QY**2 - QZ**2
    equilibrium = [
        RHO, -11 * RHO + 19 * j2, 3 * RHO - sp.Rational(11, 2) * j2,
        QX, -sp.Rational(2, 3) * QX, QY, -sp.Rational(2, 3) * QY, QZ, -sp.Rational(2, 3) * QZ,
        pxx, -sp.Rational(1, 2) * pxx, pww, -sp.Rational(1, 2) * pww,
        QX * QY, QY * QZ, QZ * QX, 0, 0, 0,
    ]
    s9 = 1.0 / (3.0 * nu + 0.5)   # nu in lattice units
    s = [0.0, 1.19, 1.4, 0.0, 1.2, 0.0, 1.2, 0.0, 1.2, s9, 1.4, s9, 1.4, s9, s9, s9, 1.98, 1.98, 1.98]
    init = {RHO: rho_o, QX: 0.0, QY: 0.0, QZ: 0.0}
    if perturb is not None:
        init = {
            RHO: _perturbed(perturb, rho_o), QX: _perturbed(perturb + 1, 0.0),
            QY: _perturbed(perturb + 2, 0.0), QZ: _perturbed(perturb + 3, 0.0),
        }
    return {
        "box": {"x": [0.0, 1.0], "y": [0.0, 1.0], "z": [0.0, 1.0], "label": [0, 0, 0, 0, 0, 1]},
        "space_step": dx,
        "scheme_velocity": 1.0,
        "schemes": [
            {
                "velocities": list(range(19)),
                "polynomials": polynomials,
                "relaxation_parameters": s,
                "equilibrium": equilibrium,
                "conserved_moments": [RHO, QX, QY, QZ],
            }
        ],
        "parameters": {LA: 1.0},
        "init": init,
        "boundary_conditions": {
            0: {"method": {0: mod.bc.BounceBack}},
            1: {"method": {0: mod.bc.BounceBack}, "value": (_wall_value_3d, (rho_o, lid))},
        },
        "generator": generator,
    }


# --------------------------------------------------------------------------
# D3Q27 channel with a sphere
# --------------------------------------------------------------------------
def channel_sphere_d3q27(nx=256, ny=128, nz=128, mod=None, perturb=None, generator="cuda", nu=0.01,
                         relative_velocity=False):
    """C5: D3Q27 tensor-product raw moments, Poiseuille-like channel (inlet bounce-back with a
    prescribed velocity, Neumann outlet, bounce-back walls) around a Bouzidi sphere."""
    mod = mod or _default_mod()
    dx = 1.0 / ny
    lx, ly, lz = nx * dx, 1.0, nz * dx
    rho_o, u_o = 1.0, 0.04
    orders = sorted(((a, b, c) for a in range(3) for b in range(3) for c in range(3)),
                    key=lambda t: (sum(t), t[::-1]))
    # conserved rows first in a readable order: mass, X, Y, Z
    head = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    orders = head + [o for o in orders if o not in head]
    ux, uy, uz = QX / RHO, QY / RHO, QZ / RHO

    def mu(k, u):
        return [1, u, u**2 + sp.Rational(1, 3)][k]

    polynomials = [X**a * Y**b * Z**c for a, b, c in orders]
    equilibrium = [RHO, QX, QY, QZ] + [RHO * mu(a, ux) * mu(b, uy) * mu(c, uz) for a, b, c in orders[4:]]
    s_nu = 1.0 / (3.0 * nu + 0.5)
    s = [0.0 if sum(o) <= 1 else (s_nu if sum(o) == 2 else 1.0) for o in orders]
    init = {RHO: rho_o, QX: rho_o * u_o, QY: 0.0, QZ: 0.0}
    if perturb is not None:
        init = {
            RHO: _perturbed(perturb, rho_o), QX: _perturbed(perturb + 1, rho_o * u_o),
            QY: _perturbed(perturb + 2, 0.0), QZ: _perturbed(perturb + 3, 0.0),
        }
    dico = {
        "box": {"x": [0.0, lx], "y": [0.0, ly], "z": [0.0, lz], "label": [1, 2, 0, 0, 0, 0]},
        "elements": [mod.Sphere([0.25 * lx, 0.5 * ly + dx, 0.5 * lz], 0.125, label=3)],
        "space_step": dx,
        "scheme_velocity": 1.0,
        "schemes": [
            {
                "velocities": list(range(27)),
                "polynomials": polynomials,
                "relaxation_parameters": s,
                "equilibrium": equilibrium,
                "conserved_moments": [RHO, QX, QY, QZ],
            }
        ],
        "parameters": {LA: 1.0},
        "init": init,
        "boundary_conditions": {
            0: {"method": {0: mod.bc.BounceBack}},
            1: {"method": {0: mod.bc.BounceBack}, "value": (_wall_value_3d, (rho_o, u_o))},
            2: {"method": {0: mod.bc.NeumannX}},
            3: {"method": {0: mod.bc.BouzidiBounceBack}},
        },
        "generator": generator,
    }
    if relative_velocity:
        dico["relative_velocity"] = [QX / RHO, QY / RHO, QZ / RHO]
    return dico


CASES = {
    "lid_cavity_d2q9": lid_cavity_d2q9,
    "karman_d2q9": karman_d2q9,
    "shallow_water_d2q4": shallow_water_d2q4,
    "lid_cavity_d3q19": lid_cavity_d3q19,
    "channel_sphere_d3q27": channel_sphere_d3q27,
}


# --------------------------------------------------------------------------
# additional parity workloads (features of the reference demos not covered by C1-C5)
# --------------------------------------------------------------------------
T = sp.symbols("T")
U = sp.symbols("u")


def _rb_init_T(x, y, Td, Tu, xmin, xmax, ymin, ymax):
    xmid, ymid = (xmax + xmin) / 2, (ymax + ymin) / 2
    return Td + (Tu - Td) / (ymax - ymin) * (y - ymin) + Td * (x < 1.1 * xmid) * (x > 0.9 * xmid) * (y < ymid)


def _rb_up(f, m, x, y, Tu):
    m[QX] = 0.0
    m[QY] = 0.0
    m[T] = Tu


def _rb_down(f, m, x, y, Td):
    m[QX] = 0.0
    m[QY] = 0.0
    m[T] = Td


def _rb_down_with_time(f, m, t, x, y, Td, Tu):
    m[QX] = 0.0
    m[QY] = 0.0
    if 0 <= t % 10.0 <= 5:
        m[T] = Td
    else:
        m[T] = Tu


def rayleigh_benard(nx=128, ny=64, mod=None, perturb=None, generator="cuda", time_bc=True, period=10.0):
    """D2Q9 (fluid, source term alpha*g*T on qy) coupled to D2Q5 (temperature): Bouzidi bounce-back
    on the fluid and Bouzidi ANTI bounce-back on the temperature, periodic in x, bottom value that
    depends on time (demo/2D/Rayleigh-Benard.py:62-232; reference golden test2D_rayleigh_benard.h5
    is nx=128, ny=64, Tf=0.5, time_bc=True)."""
    mod = mod or _default_mod()
    Td, Tu = 0.5, -0.5
    xmin, xmax, ymin, ymax = 0.0, 2.0, 0.0, 1.0
    dx = (ymax - ymin) / ny
    Ra, Pr, alpha, la, g = 2000, 0.71, 0.005, 1.0, 9.81
    nu = np.sqrt(Pr * alpha * 9.81 * (Td - Tu) * (ymax - ymin) / Ra)
    kappa = nu / Pr
    eta = nu
    snu = 1.0 / (0.5 + 3 * nu)
    seta = 1.0 / (0.5 + 3 * eta)
    sq = 8 * (2 - snu) / (8 - snu)
    sf = [0.0, 0.0, 0.0, seta, seta, sq, sq, snu, snu]
    a = 0.5
    skappa = 1.0 / (0.5 + 10 * kappa / (4 + a))
    se = 1.0 / (0.5 + np.sqrt(3) / 3)
    sT = [0.0, skappa, skappa, se, se]

    if period != 10.0:
        def down_t(f, m, t, x, y, Td, Tu):
            m[QX] = 0.0
            m[QY] = 0.0
            m[T] = Td if 0 <= t % period <= period / 2 else Tu
    else:
        down_t = _rb_down_with_time
    methods = {0: mod.bc.BouzidiBounceBack, 1: mod.bc.BouzidiAntiBounceBack}
    if time_bc:
        bottom = {"method": dict(methods), "value": (down_t, (Td, Tu)), "time_bc": True}
    else:
        bottom = {"method": dict(methods), "value": (_rb_down, (Td,))}
    init = {RHO: 1.0, QX: 0.0, QY: 0.0, T: (_rb_init_T, (Td, Tu, xmin, xmax, ymin, ymax))}
    if perturb is not None:
        init[QX] = _perturbed(perturb + 1, 0.0)
        init[QY] = _perturbed(perturb + 2, 0.0)
    return {
        "box": {"x": [xmin, xmin + nx * dx], "y": [ymin, ymax], "label": [-1, -1, 0, 1]},
        "space_step": dx,
        "scheme_velocity": la,
        "schemes": [
            {
                "velocities": list(range(9)),
                "conserved_moments": [RHO, QX, QY],
                "polynomials": [
                    1, X, Y, 3 * (X**2 + Y**2) - 4,
                    0.5 * (9 * (X**2 + Y**2) ** 2 - 21 * (X**2 + Y**2) + 8),
                    3 * X * (X**2 + Y**2) - 5 * X, 3 * Y * (X**2 + Y**2) - 5 * Y, X**2 - Y**2, X * Y,
                ],
                "relaxation_parameters": sf,
                "equilibrium": [
                    RHO, QX, QY, -2 * RHO + 3 * (QX**2 + QY**2), RHO - 3 * (QX**2 + QY**2),
                    -QX, -QY, QX**2 - QY**2, QX * QY,
                ],
                "source_terms": {QY: alpha * g * T},
            },
            {
                "velocities": list(range(5)),
                "conserved_moments": T,
                "polynomials": [1, X, Y, 5 * (X**2 + Y**2) - 4, (X**2 - Y**2)],
                "equilibrium": [T, T * QX, T * QY, a * T, 0.0],
                "relaxation_parameters": sT,
            },
        ],
        "init": init,
        "boundary_conditions": {0: bottom, 1: {"method": dict(methods), "value": (_rb_up, (Tu,))}},
        "generator": generator,
    }


def _bump(x, xmin, xmax):
    mid, width = 0.5 * (xmin + xmax), 0.125 * (xmax - xmin)
    return np.exp(-((x - mid) / width) ** 2)


def _u_left(f, m, x, value):
    m[U] = value


def advection_d1q5(n=128, mod=None, perturb=None, generator="cuda"):
    """1-D advection, D1Q5 (velocities up to +-2: two ghost layers per side), Neumann outflow on the
    right and a prescribed value with bounce-back on the left."""
    mod = mod or _default_mod()
    xmin, xmax = 0.0, 1.0
    dx = (xmax - xmin) / n
    c = 0.4
    init = {U: (_bump, (xmin, xmax))}
    if perturb is not None:
        init = {U: _perturbed(perturb, 0.5, amp=0.2)}
    return {
        "box": {"x": [xmin, xmax], "label": [0, 1]},
        "space_step": dx,
        "scheme_velocity": 1.0,
        "schemes": [
            {
                "velocities": list(range(5)),
                "conserved_moments": U,
                "polynomials": [1, X, X**2 / 2, X**3 / 6, X**4 / 24],
                "equilibrium": [U, c * U, c**2 * U / 2 + U / 6, c**3 * U / 6, c**4 * U / 24],
                "relaxation_parameters": [0.0, 1.5, 1.2, 1.0, 1.3],
            }
        ],
        "init": init,
        "boundary_conditions": {
            0: {"method": {0: mod.bc.BounceBack}, "value": (_u_left, (0.25,))},
            1: {"method": {0: mod.bc.Neumann}},
        },
        "generator": generator,
    }


def _t_wall(f, m, x, y, value):
    m[T] = value


def heat_d2q5(n=48, mod=None, perturb=None, generator="cuda", plain=False):
    """2-D heat equation, D2Q5 with anti bounce-back (Dirichlet) walls, a solid triangle and an
    elliptic hole treated with Bouzidi anti bounce-back, Neumann in y on the top wall.
    plain=True: periodic in x, anti bounce-back (two different temperatures) on both y walls, no
    obstacle -- the 2-D case whose walls the fused kernel can apply itself (lbmk_walls)."""
    mod = mod or _default_mod()
    dx = 1.0 / n
    init = {T: 0.0}
    if perturb is not None:
        init = {T: _perturbed(perturb, 0.5, amp=0.3)}
    if plain:
        return {
            "box": {"x": [0.0, 1.0], "y": [0.0, 1.0], "label": [-1, -1, 0, 1]},
            "space_step": dx,
            "scheme_velocity": 1.0,
            "schemes": [
                {
                    "velocities": list(range(5)),
                    "conserved_moments": T,
                    "polynomials": [1, X, Y, (X**2 + Y**2) / 2, (X**2 - Y**2) / 2],
                    "equilibrium": [T, 0.0, 0.0, 0.4 * T, 0.0],
                    "relaxation_parameters": [0.0, 1.2, 1.2, 1.5, 1.1],
                }
            ],
            "init": init,
            "boundary_conditions": {
                0: {"method": {0: mod.bc.AntiBounceBack}, "value": (_t_wall, (1.0,))},
                1: {"method": {0: mod.bc.AntiBounceBack}, "value": (_t_wall, (0.25,))},
            },
            "generator": generator,
        }
    return {
        "box": {"x": [0.0, 1.0], "y": [0.0, 1.0], "label": [0, 1, 2, 3]},
        "elements": [
            mod.Triangle([0.15, 0.2], [0.3, 0.05], [0.05, 0.3], label=4),
            mod.Ellipse([0.65, 0.6], [0.2, 0.1], [-0.05, 0.1], label=5),
        ],
        "space_step": dx,
        "scheme_velocity": 1.0,
        "schemes": [
            {
                "velocities": list(range(5)),
                "conserved_moments": T,
                "polynomials": [1, X, Y, (X**2 + Y**2) / 2, (X**2 - Y**2) / 2],
                "equilibrium": [T, 0.0, 0.0, 0.4 * T, 0.0],
                "relaxation_parameters": [0.0, 1.2, 1.2, 1.5, 1.1],
            }
        ],
        "init": init,
        "boundary_conditions": {
            0: {"method": {0: mod.bc.AntiBounceBack}, "value": (_t_wall, (1.0,))},
            1: {"method": {0: mod.bc.AntiBounceBack}, "value": (_t_wall, (0.0,))},
            2: {"method": {0: mod.bc.AntiBounceBack}, "value": (_t_wall, (0.5,))},
            3: {"method": {0: mod.bc.NeumannY}},
            4: {"method": {0: mod.bc.BouzidiAntiBounceBack}, "value": (_t_wall, (0.25,))},
            5: {"method": {0: mod.bc.BouzidiAntiBounceBack}, "value": (_t_wall, (0.75,))},
        },
        "generator": generator,
    }


def advection_d3q6(n=16, mod=None, perturb=None, generator="cuda"):
    """3-D advection, D3Q6, fully periodic, initialisation on the DISTRIBUTIONS (inittype)."""
    dx = 1.0 / n
    cx, cy, cz = 0.2, 0.1, -0.1

    def blob(x, y, z):
        return 1.0 + np.exp(-40 * ((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2))

    if perturb is not None:
        finit = {k: _perturbed(perturb + k, 1.0 / 6, amp=0.05) for k in range(6)}
    else:
        finit = {k: (lambda x, y, z: blob(x, y, z) / 6) for k in range(6)}
    return {
        "box": {"x": [0.0, 1.0], "y": [0.0, 1.0], "z": [0.0, 1.0], "label": -1},
        "space_step": dx,
        "scheme_velocity": 1.0,
        "schemes": [
            {
                "velocities": list(range(1, 7)),
                "conserved_moments": U,
                "polynomials": [1, X, Y, Z, X**2 - Y**2, X**2 - Z**2],
                "equilibrium": [U, cx * U, cy * U, cz * U, 0.0, 0.0],
                "relaxation_parameters": [0.0, 1.6, 1.6, 1.6, 1.3, 1.3],
            }
        ],
        "inittype": "distributions",
        "init": finit,
        "generator": generator,
    }


def advection_d2q13(nx=24, ny=20, mod=None, perturb=None, generator="cuda"):
    """2-D advection-diffusion on D2Q13 (velocities up to +-2 on the axes: TWO ghost layers per side in
    both directions), ragged sizes, periodic in y, prescribed value (bounce-back) on the left and
    Neumann outflow on the right."""
    mod = mod or _default_mod()
    dx = 1.0 / ny
    cx, cy = 0.3, -0.15
    pol = [1, X, Y, X**2 + Y**2, X**2 - Y**2, X * Y, X**2 * Y, X * Y**2, X**3, Y**3,
           X**2 * Y**2, X**4 + Y**4, X**4 - Y**4]
    eq = [U, cx * U, cy * U, (cx**2 + cy**2 + 1.0) * U, (cx**2 - cy**2) * U, cx * cy * U,
          cx * U / 2, cy * U / 2, cx * U, cy * U, U / 3, 2 * U, 0.0]
    s = [0.0, 1.3, 1.3, 1.1, 1.5, 1.5, 1.2, 1.2, 1.4, 1.4, 1.0, 1.6, 1.6]

    def blob(x, y):
        return 1.0 + np.exp(-30 * ((x - 0.5) ** 2 + (y - 0.5) ** 2))

    init = {U: blob}
    if perturb is not None:
        init = {U: _perturbed(perturb, 1.0, amp=0.2)}
    return {
        "box": {"x": [0.0, nx * dx], "y": [0.0, 1.0], "label": [0, 1, -1, -1]},
        "space_step": dx,
        "scheme_velocity": 1.0,
        "schemes": [
            {"velocities": list(range(13)), "conserved_moments": U, "polynomials": pol,
             "equilibrium": eq, "relaxation_parameters": s}
        ],
        "init": init,
        "boundary_conditions": {
            0: {"method": {0: mod.bc.BounceBack}, "value": (_u_left_2d, (1.25,))},
            1: {"method": {0: mod.bc.NeumannX}},
        },
        "generator": generator,
    }


def _u_left_2d(f, m, x, y, value):
    m[U] = value


CASES.update({
    "advection_d2q13": advection_d2q13,
    "rayleigh_benard": rayleigh_benard,
    "advection_d1q5": advection_d1q5,
    "heat_d2q5": heat_d2q5,
    "advection_d3q6": advection_d3q6,
})
