"""
In-place streaming (AA pattern, `in_place=True`): ONE population array, even steps gather and scatter
back, odd steps are local (include/lbm_b200.h: lbm_sim_set_aa; the algorithm is proven against the
reference order by brute force in tests/test_aa_emulation.py).  On the GPU the populations of the fluid
cells must be IDENTICAL to the two-array kernel's after an odd and after an even number of steps, on the
12 parity workloads: single steps, CUDA-graph pairs, a stand-alone boundary_condition(), time-dependent
boundary values, periodic boxes, two ghost layers, obstacles.
"""
import numpy as np
import pytest

from conftest import PARITY_CASES, case_id

pytestmark = pytest.mark.gpu
IDS = [case_id(*c) for c in PARITY_CASES]


@pytest.fixture(autouse=True)
def one_lowering(monkeypatch):
    """both variants run kernels of ONE generated library (the one that also holds the in-place
    launchers): separately generated libraries of the same scheme differ in how sympy.cse groups the
    sums, i.e. in last bits; against the standard library the agreement is 1e-13 (checked below)."""
    monkeypatch.setenv("PYLBM_B200_AA_LIBRARY", "1")


def _fluid(sim):
    inner = tuple(slice(v, -v) for v in sim.domain.stencil.vmax)
    return sim.domain.in_or_out[inner] == sim.domain.valin


def _compare(a, b):
    """a: in-place simulation, b: two-array simulation, same number of steps."""
    fluid = _fluid(b)
    for k in range(b.container.nv):
        fa, fb = a.F[k], b.F[k]
        assert np.array_equal(fa[fluid], fb[fluid]), "population %d" % k
    for key in b.scheme.consm:
        ma, mb = a.m[key], b.m[key]
        np.testing.assert_allclose(ma[fluid], mb[fluid], rtol=1e-14, atol=1e-300)


@pytest.mark.parametrize("name,kw", PARITY_CASES, ids=IDS)
def test_in_place_streaming_is_bit_identical_to_two_arrays(name, kw):
    import pylbm_b200
    from pylbm_b200 import cases

    a = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), in_place=True)
    b = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw))
    assert a.container.Fnew is a.container.F and b.container.Fnew is not b.container.F
    assert (a.bc.walls is None) == (b.bc.walls is None)      # the fused walls work in place as well
    for sim in (a, b):
        for _ in range(3):
            sim.one_time_step()
    assert a._swapped
    _compare(a, b)                      # odd: read through the swapped layout
    for sim in (a, b):
        sim.run(4)                      # swapped start: one single step, a graph pair, one single step
    assert a._swapped and a.nt == 7
    _compare(a, b)
    for sim in (a, b):
        sim.boundary_condition()        # stand-alone, on the odd-step lists
        sim.one_time_step()
    assert not a._swapped
    _compare(a, b)                      # even: natural layout
    inner = (slice(None),) + tuple(slice(v, -v) for v in a.domain.stencil.vmax)
    fluid = _fluid(b)
    assert np.array_equal(a.container.F.get()[inner][:, fluid], b.container.F.get()[inner][:, fluid])
    for sim in (a, b):
        sim.boundary_condition()
        sim.run(6)
    _compare(a, b)
    with pytest.raises(NotImplementedError):
        a.transport()


@pytest.mark.parametrize("name,kw", [PARITY_CASES[i] for i in (1, 4, 5)], ids=[IDS[i] for i in (1, 4, 5)])
@pytest.mark.parametrize("compute", [None, "float32"])
def test_in_place_streaming_with_fp32_populations(name, kw, compute):
    import pylbm_b200
    from pylbm_b200 import cases

    a = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), dtype="float32", compute_dtype=compute, in_place=True)
    b = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), dtype="float32", compute_dtype=compute)
    for sim in (a, b):
        sim.run(9)
    _compare(a, b)
    for sim in (a, b):
        sim.one_time_step()
    _compare(a, b)


def test_outside_writes_need_the_natural_layout():
    import pylbm_b200
    from pylbm_b200 import cases

    sim = pylbm_b200.Simulation(cases.karman_d2q9(nx=64, ny=32, perturb=0), in_place=True)
    sim.one_time_step()
    with pytest.raises(RuntimeError):
        sim.F_halo[0] = sim.F[0]
    with pytest.raises(RuntimeError):
        sim.m2f()
    sim.one_time_step()
    f0 = sim.F_halo[0]
    sim.F_halo[0] = f0                  # even: allowed, ghosts are refreshed by the next step
    ref = pylbm_b200.Simulation(cases.karman_d2q9(nx=64, ny=32, perturb=0))
    ref.run(2)
    sim.run(3)
    ref.run(3)
    _compare(sim, ref)


def test_in_place_streaming_through_pylbm_simulation(pylbm):
    """the north-star interface: `dico['cuda_option'] = {'in_place': True}`."""
    from pylbm_b200 import cases, plugin

    plugin.register()
    d = cases.lid_cavity_d3q19(n=16, perturb=0, mod=pylbm, generator="cuda")
    d["cuda_option"] = {"in_place": True}
    a = pylbm.Simulation(d)
    b = pylbm.Simulation(cases.lid_cavity_d3q19(n=16, perturb=0, mod=pylbm, generator="cuda"))
    assert a.container.Fnew is a.container.F
    for sim in (a, b):
        for _ in range(5):
            sim.one_time_step()
    _compare(a, b)
    for sim in (a, b):
        sim.one_time_step()
    _compare(a, b)


def test_in_place_streaming_against_the_standard_library_and_the_oracle(monkeypatch):
    """in place (its own library) vs the standard two-array library vs the oracle: 1e-12 of max|field|."""
    import pylbm_b200
    from pylbm_b200 import cases
    from oracle.lbm_oracle import OracleSimulation

    monkeypatch.delenv("PYLBM_B200_AA_LIBRARY", raising=False)
    name, kw = PARITY_CASES[5]
    a = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), in_place=True)
    b = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw))
    ora = OracleSimulation(cases.CASES[name](perturb=0, **kw))
    for nsteps in (25, 1):
        a.run(nsteps)
        b.run(nsteps)
        for _ in range(nsteps):
            ora.one_time_step()
        fluid = _fluid(b)
        for key in b.scheme.consm:
            okey = [k for k in ora.scheme.consm if str(k) == str(key)][0]
            want = ora.m[okey][fluid]
            scale = np.abs(want).max()
            assert np.abs(a.m[key][fluid] - want).max() <= 1e-12 * scale
            assert np.abs(a.m[key][fluid] - b.m[key][fluid]).max() <= 1e-13 * scale


@pytest.mark.parametrize("name,kw", [("lid_cavity_d3q19", dict(n=16)), ("channel_sphere_d3q27", dict(nx=21, ny=13, nz=9)),
                                     ("heat_d2q5", dict(n=24, plain=True))])
def test_in_place_streaming_with_and_without_the_fused_walls(name, kw, monkeypatch):
    """the wall plan (bounce-back walls of the fastest axis applied by the kernels) in place: the even step
    stores the bounced value into the cell's own slot; populations identical with the plan on and off,
    also after an outside write (one step through the stale-only entries)."""
    import pylbm_b200
    from pylbm_b200 import cases

    def run(walls):
        if walls:
            monkeypatch.delenv("PYLBM_B200_NO_WALLS", raising=False)
        else:
            monkeypatch.setenv("PYLBM_B200_NO_WALLS", "1")
        sim = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), in_place=True)
        assert (sim.bc.walls is not None) == walls
        sim.run(7)
        sim.one_time_step()
        sim.boundary_condition()
        sim.F_halo[1] = sim.F_halo[1]
        sim.run(5)
        return sim

    a, b = run(True), run(False)
    assert a._swapped and b._swapped
    fluid = _fluid(b)
    for k in range(b.container.nv):
        assert np.array_equal(a.F[k][fluid], b.F[k][fluid]), "population %d" % k
