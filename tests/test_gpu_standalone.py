"""
The stand-alone kernels of the path -- `transport`, `f2m`, `source_term`, `relaxation`, `m2f`,
`equilibrium`, `boundary_condition` (reference: simulation.py:322-390, algorithm/base.py:289-504) -- on the
CUDA backend against the UNMODIFIED reference run live in the same process (oracle/_ref, NumPy generator:
no build step), call by call on the same seeded perturbed state.  After every call the array the call
modifies is compared on the interior cells (ghost entries no later load reads are scratch on both sides).  Tolerance 1e-12 of
max|array| (fp64; the two sides group the sums differently).

`transport()` leaves its result in F, like the reference's NumPy backend and the reference's docstrings
("the array _F is modified", the split step of simulation.py:392-408); the reference's Cython backend
leaves it in Fnew and never swaps -- a difference between the reference's own backends.
`source_term()` integrates half a time step (ode.py:11-16 `lhs + dt/2*rhs`) whatever
`fraction_of_time_step` says: the reference does not forward that argument either (simulation.py:340-345).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [
    ("rayleigh_benard", dict(nx=48, ny=24, period=0.2), True),     # two schemes, source term, time_bc
    ("lid_cavity_d3q19", dict(n=10), False),
    ("karman_d2q9", dict(nx=48, ny=24), False),                     # relative velocity (fused kernel only)
    ("shallow_water_d2q4", dict(n=24), False),                      # vectorial, 1/h equilibria
]


def _inner(sol):
    return (slice(None),) + tuple(slice(v, -v) for v in sol.domain.stencil.vmax)


def _ref_F(ref):
    return np.ascontiguousarray(ref.container.F.swaparray)


def _ref_m(ref):
    return np.ascontiguousarray(ref.container.m.swaparray)


def _close(got, want, what):
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)
    assert err <= TOL, (what, err)


@pytest.mark.parametrize("name,kw,source", CASES, ids=[c[0] for c in CASES])
def test_standalone_kernels_against_the_reference(pylbm, name, kw, source):
    from pylbm_b200 import cases, plugin

    plugin.register()
    ref = pylbm.Simulation(cases.CASES[name](perturb=5, mod=pylbm, generator="numpy", **kw))
    sol = pylbm.Simulation(cases.CASES[name](perturb=5, mod=pylbm, generator="cuda", **kw))
    inner = _inner(sol)
    _close(sol.container.F.get()[inner], _ref_F(ref)[inner], "initial F")
    for sweep in range(3):
        for call in ("boundary_condition", "transport", "f2m", "source_term", "relaxation", "source_term", "m2f"):
            if call == "source_term" and not source:
                continue
            getattr(ref, call)()
            getattr(sol, call)()
            if call in ("boundary_condition", "transport", "m2f"):
                _close(sol.container.F.get()[inner], _ref_F(ref)[inner], (sweep, call, "F"))
            else:
                _close(sol.container.m.get()[inner], _ref_m(ref)[inner], (sweep, call, "m"))
    # equilibrium on the whole moment array
    ref.equilibrium()
    sol.equilibrium()
    _close(sol.container.m.get()[inner], _ref_m(ref)[inner], "equilibrium")
    # and the split sweeps moved the state like fused steps do (no relative velocity, no source term:
    # simulation.py:392-408)
    if name == "lid_cavity_d3q19":
        fused = pylbm.Simulation(cases.CASES[name](perturb=5, mod=pylbm, generator="cuda", **kw))
        for _ in range(3):
            fused.one_time_step()
        sol.m2f()
        ref_state = pylbm.Simulation(cases.CASES[name](perturb=5, mod=pylbm, generator="cuda", **kw))
        for _ in range(3):
            ref_state.boundary_condition()
            for call in ("transport", "f2m", "relaxation", "m2f"):
                getattr(ref_state, call)()
        _close(ref_state.container.F.get()[inner], fused.container.F.get()[inner], "split == fused")


def test_source_term_needs_a_source(pylbm):
    from pylbm_b200 import cases, plugin

    plugin.register()
    sol = pylbm.Simulation(cases.lid_cavity_d3q19(n=8, mod=pylbm, generator="cuda"))
    with pytest.raises((KeyError, AttributeError)):
        sol.source_term()
