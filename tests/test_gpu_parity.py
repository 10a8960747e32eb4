"""
Parity of the CUDA path against the oracle (tests/ may import oracle/; the product may not).

fp64 acceptance: conserved moments on fluid cells after N steps,
max|delta| / max|ref| <= 1e-12 (BASELINE.json north_star); boundary index lists bit-exact.
"""
import numpy as np
import pytest

from conftest import PARITY_CASES

pytestmark = pytest.mark.gpu

TOL_F64 = 1e-12
TOL_F32_STORAGE = 5e-5   # fp32 storage of the populations, fp64 arithmetic, 50 steps (measured <= 2.1e-5)
TOL_F32_ARITHMETIC = 2e-4   # fp32 storage and fp32 arithmetic in the time-step kernel, 50 steps (measured <= 6.9e-5)


def _build(name, kw, perturb=0, **simkw):
    import pylbm_b200
    from pylbm_b200 import cases
    from oracle.lbm_oracle import OracleSimulation

    dico = cases.CASES[name](perturb=perturb, **kw)
    sim = pylbm_b200.Simulation(dico, **simkw)
    ora = OracleSimulation(cases.CASES[name](perturb=perturb, **kw))
    return sim, ora


def _compare(sim, ora, tol):
    fluid = ora.domain.in_or_out[tuple(slice(v, -v) for v in ora.domain.stencil.vmax)] == ora.domain.valin
    worst = 0.0
    for key in sim.scheme.consm:
        okey = [k for k in ora.scheme.consm if str(k) == str(key)][0]
        a, b = sim.m[key], ora.m[okey]
        assert a.shape == b.shape
        scale = np.abs(b[fluid]).max()
        err = np.abs(a[fluid] - b[fluid]).max() / max(scale, 1e-300)
        worst = max(worst, err)
    assert worst <= tol, "relative error %.3e > %.1e" % (worst, tol)
    return worst


@pytest.mark.parametrize("name,kw", PARITY_CASES, ids=[c[0] + "-" + "x".join(str(v) for v in c[1].values()) for c in PARITY_CASES])
def test_lists_and_moments(name, kw):
    sim, ora = _build(name, kw)
    # boundary lists: identical to the oracle's (both pinned against the reference fixtures)
    assert len(sim.bc.methods) == len(ora.bc.methods)
    for a, b in zip(sim.bc.methods, ora.bc.methods):
        assert type(a) is type(b)
        assert np.array_equal(a.istore, b.istore)
        for x, y in zip(a.iload, b.iload):
            assert np.array_equal(x, y)
        np.testing.assert_allclose(a.rhs, b.rhs, rtol=0, atol=1e-15)
    # initial state
    _compare(sim, ora, 1e-13)
    for _ in range(50):
        sim.one_time_step()
        ora.one_time_step()
    _compare(sim, ora, TOL_F64)
    # populations including ghosts of the current array: same state, not only same moments
    F = sim.container.F.get()
    Fo = ora._F.swaparray
    inner = (slice(None),) + tuple(slice(v, -v) for v in ora.domain.stencil.vmax)
    assert np.abs(F[inner] - Fo[inner]).max() <= 1e-12 * max(1.0, np.abs(Fo).max())


@pytest.mark.parametrize("name,kw", [PARITY_CASES[1], PARITY_CASES[4]])
def test_run_many_steps_matches_step_by_step(name, kw):
    """run(n) (CUDA-graph pairs, one runtime call) == n x one_time_step()."""
    sim_a, _ = _build(name, kw)
    sim_b, _ = _build(name, kw)
    for _ in range(21):
        sim_a.one_time_step()
    sim_b.run(21)
    assert sim_a.nt == sim_b.nt == 21
    for key in sim_a.scheme.consm:
        assert np.array_equal(sim_a.m[key], sim_b.m[key])


@pytest.mark.parametrize("name,kw", [PARITY_CASES[i] for i in (0, 1, 3, 4, 5, 6)],
                         ids=[PARITY_CASES[i][0] for i in (0, 1, 3, 4, 5, 6)])
def test_fp32_storage_mode(name, kw):
    sim, ora = _build(name, kw, dtype="float32")
    for _ in range(50):
        sim.one_time_step()
        ora.one_time_step()
    _compare(sim, ora, TOL_F32_STORAGE)


@pytest.mark.parametrize("name,kw", [PARITY_CASES[i] for i in (0, 1, 3, 4, 5, 6)],
                         ids=[PARITY_CASES[i][0] for i in (0, 1, 3, 4, 5, 6)])
def test_fp32_arithmetic_mode(name, kw):
    """all-single-precision mode: fp32 storage AND fp32 arithmetic in the time-step kernel (moments are
    still evaluated in fp64 from the stored populations).  Stated tolerance: 2e-4 relative to
    max|field| after 50 steps (measured: see DESIGN.md)."""
    sim, ora = _build(name, kw, dtype="float32", compute_dtype="float32")
    for _ in range(50):
        sim.one_time_step()
        ora.one_time_step()
    worst = _compare(sim, ora, TOL_F32_ARITHMETIC)
    print("fp32 arithmetic", name, kw, "max rel err", worst)


def test_mass_conservation_periodic():
    """size-independent property: a fully periodic box conserves every conserved moment."""
    import pylbm_b200
    from pylbm_b200 import cases

    sim = pylbm_b200.Simulation(cases.shallow_water_d2q4(n=256, perturb=3))
    before = {k: sim.m[k].sum() for k in sim.scheme.consm}
    sim.run(100)
    for k, v in before.items():
        after = sim.m[k].sum()
        assert abs(after - v) <= 1e-10 * max(1.0, abs(v))


def test_boundary_condition_only():
    """sol.boundary_condition() (ghost update + boundary kernels) against the oracle.

    Interior cells (obstacle cells included: boundary kernels store there) must agree after the call;
    ghost cells are scratch in both implementations except for the entries a later pull or boundary
    load reads, which the following step exercises."""
    name, kw = PARITY_CASES[1]
    sim, ora = _build(name, kw)
    for _ in range(3):
        sim.one_time_step()
        ora.one_time_step()
    sim.boundary_condition()
    ora.boundary_condition()
    inner = (slice(None),) + tuple(slice(v, -v) for v in ora.domain.stencil.vmax)
    assert np.abs(sim.container.F.get()[inner] - ora._F.swaparray[inner]).max() <= 1e-13
    # the stored ghost entries that matter: populations entering the domain
    F, Fo = sim.container.F.get(), ora._F.swaparray
    vel = sim.scheme.stencil.get_all_velocities()
    for k, v in enumerate(vel):
        if v[0] > 0:
            assert np.abs(F[k, 0, 1:-1] - Fo[k, 0, 1:-1]).max() <= 1e-13
        if v[0] < 0:
            assert np.abs(F[k, -1, 1:-1] - Fo[k, -1, 1:-1]).max() <= 1e-13
        if v[1] > 0:
            assert np.abs(F[k, 1:-1, 0] - Fo[k, 1:-1, 0]).max() <= 1e-13
        if v[1] < 0:
            assert np.abs(F[k, 1:-1, -1] - Fo[k, 1:-1, -1]).max() <= 1e-13
    sim.one_time_step()
    ora.one_time_step()
    assert np.abs(sim.container.F.get()[inner] - ora._F.swaparray[inner]).max() <= 1e-13


# ---------------------------------------------------------------------------
# directly against the REFERENCE fixtures (tests/golden, tools/make_golden.py)
# ---------------------------------------------------------------------------
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,kw", PARITY_CASES, ids=[c[0] + "-" + "x".join(str(v) for v in c[1].values()) for c in PARITY_CASES])
def test_cuda_against_reference_fixture(name, kw):
    """boundary lists bit-exact, conserved moments after 50 steps within 1e-12 of the reference's
    Cython generator (fixture produced by the unmodified pylbm)."""
    import pylbm_b200
    from pylbm_b200 import cases
    from conftest import case_id

    ref = np.load(os.path.join(GOLDEN, "ref_%s.npz" % case_id(name, kw)))
    sim = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw))
    assert len(sim.bc.methods) == int(ref["nmethods"])
    for i, method in enumerate(sim.bc.methods):
        pre = "bc%d_" % i
        assert type(method).__name__ == str(ref[pre + "name"])
        assert method.istore.dtype == np.int32 and np.array_equal(method.istore, ref[pre + "istore"])
        for j, il in enumerate(method.iload):
            assert np.array_equal(il, ref[pre + "iload%d" % j])
        if hasattr(method, "s"):
            assert np.array_equal(method.s, ref[pre + "s"])
        np.testing.assert_allclose(method.rhs, ref[pre + "rhs"], rtol=0, atol=1e-15)
    sim.run(int(ref["nsteps"]))
    fluid = ref["in_or_out"][tuple(slice(v, -v) for v in sim.domain.stencil.vmax)] == sim.domain.valin
    for key in sim.scheme.consm:
        a, b = sim.m[key], ref["m_" + str(key)]
        err = np.abs(a[fluid] - b[fluid]).max() / np.abs(b[fluid]).max()
        assert err <= TOL_F64, (str(key), err)


def test_cuda_against_reference_golden_h5_fields():
    """the reference's own golden fields (dx = 1/64, Tf = 0.5, solid cells zeroed)."""
    import pylbm_b200
    from pylbm_b200 import cases

    manifest = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))
    for fname, meta in manifest["files"].items():
        if not fname.startswith("h5_"):
            continue
        ref = np.load(os.path.join(GOLDEN, fname))
        sim = pylbm_b200.Simulation(cases.CASES[meta["case"]](**meta["kwargs"]))
        while sim.t < meta["final_time"]:
            sim.one_time_step()
        solid = sim.domain.in_or_out[tuple(slice(v, -v) for v in sim.domain.stencil.vmax)] != sim.domain.valin
        for key in sim.scheme.consm:
            field = sim.m[key].copy()
            field[solid] = 0
            assert np.abs(field - ref[str(key)]).max() <= 1e-12, (fname, str(key))


def test_split_step_equals_fused_step():
    """boundary_condition -> transport -> f2m -> relaxation -> m2f (the stand-alone kernels, reference
    simulation.py:392-408 docstring) reproduces one_time_step for a scheme without relative velocity."""
    import pylbm_b200
    from pylbm_b200 import cases

    kw = dict(n=12)
    a = pylbm_b200.Simulation(cases.lid_cavity_d3q19(perturb=1, **kw))
    b = pylbm_b200.Simulation(cases.lid_cavity_d3q19(perturb=1, **kw))
    for _ in range(3):
        a.one_time_step()
        b.boundary_condition()
        b.transport()
        b.f2m()
        b.relaxation()
        b.m2f()
    inner = (slice(None),) + (slice(1, -1),) * 3
    Fa, Fb = a.container.F.get()[inner], b.container.F.get()[inner]
    assert np.abs(Fa - Fb).max() <= 1e-13 * np.abs(Fa).max()


def test_runtime_scalars_from_extra_parameters():
    """a symbol left free in the dictionary is a runtime scalar of the kernels, given through
    sol.extra_parameters (reference: algorithm/base.py:662-667)."""
    import sympy as sp
    import pylbm_b200
    from pylbm_b200 import cases

    grav = sp.Symbol("gravity_coeff")
    d_num = cases.rayleigh_benard(nx=32, ny=16, time_bc=False)
    d_sym = cases.rayleigh_benard(nx=32, ny=16, time_bc=False)
    num = d_num["schemes"][0]["source_terms"][cases.QY]          # alpha * g * T  (numbers)
    coeff = float(num.coeff(cases.T))
    d_sym["schemes"][0]["source_terms"] = {cases.QY: grav * cases.T}
    a = pylbm_b200.Simulation(d_num)
    b = pylbm_b200.Simulation(d_sym)
    assert "gravity_coeff" in b.kernels.scalars("one_time_step")
    b.extra_parameters[grav] = coeff
    for _ in range(10):
        a.one_time_step()
        b.one_time_step()
    for key in a.scheme.consm:
        assert np.abs(a.m[key] - b.m[key]).max() <= 1e-14
    # m_halo / F_halo setters
    rho = a.m_halo[cases.RHO]
    assert rho.shape == tuple(a.domain.shape_halo)
    f0 = a.F_halo[0]
    a.F_halo[0] = f0 * 1.0
    a.one_time_step()
    b.one_time_step()
    for key in a.scheme.consm:
        assert np.abs(a.m[key] - b.m[key]).max() <= 1e-14


@pytest.mark.parametrize("name,kw,dtype", [
    ("lid_cavity_d3q19", dict(n=16), "float64"), ("channel_sphere_d3q27", dict(nx=21, ny=13, nz=9), "float64"),
    ("lid_cavity_d3q19", dict(n=12), "float32"), ("heat_d2q5", dict(n=24, plain=True), "float64"),
])
def test_walls_in_the_fused_kernel_are_bit_identical_to_the_list_kernel(name, kw, dtype, monkeypatch):
    """bounce-back walls normal to the fastest axis are applied by the fused kernel (lbmk_walls) when
    boundary.plan_walls proves it equivalent; the arithmetic is the list kernel's, so the populations
    must be IDENTICAL to a run with the plan switched off -- including after an outside write of F,
    which sends one step through the stale-only fallback entries."""
    import pylbm_b200
    from pylbm_b200 import cases

    monkeypatch.setenv("PYLBM_B200_TASKS", "0")        # (small lattices would take the task table instead)

    def run(walls):
        if walls:
            monkeypatch.delenv("PYLBM_B200_NO_WALLS", raising=False)
        else:
            monkeypatch.setenv("PYLBM_B200_NO_WALLS", "1")
        sim = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), dtype=dtype)
        assert (sim.bc.walls is not None) == walls
        sim.run(9)
        sim.one_time_step()
        sim.boundary_condition()
        sim.F_halo[1] = sim.F_halo[1]          # outside write: ghosts stale, next step uses the full lists
        sim.run(6)
        sim.one_time_step()
        return sim

    a, b = run(True), run(False)
    replaced = sum(int(m._wall_mask.sum()) for m in a.bc.methods if getattr(m, "_wall_mask", None) is not None)
    assert replaced > 0
    inner = (slice(None),) + tuple(slice(v, -v) for v in a.domain.stencil.vmax)
    assert np.array_equal(a.container.F.get()[inner], b.container.F.get()[inner])
    for key in a.scheme.consm:
        assert np.array_equal(a.m[key], b.m[key])
    if dtype == "float64":
        # and both agree with the oracle
        from oracle.lbm_oracle import OracleSimulation

        ora = OracleSimulation(cases.CASES[name](perturb=0, **kw))
        for _ in range(17):
            ora.one_time_step()
        fluid = ora.domain.in_or_out[tuple(slice(v, -v) for v in ora.domain.stencil.vmax)] == ora.domain.valin
        for key in a.scheme.consm:
            okey = [k for k in ora.scheme.consm if str(k) == str(key)][0]
            err = np.abs(a.m[key][fluid] - ora.m[okey][fluid]).max() / np.abs(ora.m[okey][fluid]).max()
            assert err <= TOL_F64, (str(key), err)


def test_wall_plan_is_refused_when_it_would_change_results(monkeypatch):
    """Bouzidi walls, Neumann faces or periodic boxes never get the fused-kernel walls."""
    import pylbm_b200
    from pylbm_b200 import cases

    monkeypatch.setenv("PYLBM_B200_TASKS", "0")

    for name, kw in [("karman_d2q9", dict(nx=64, ny=32)), ("shallow_water_d2q4", dict(n=32)), ("heat_d2q5", dict(n=32))]:
        sim = pylbm_b200.Simulation(cases.CASES[name](**kw))
        assert sim.bc.walls is None, name


TASK_CASES = [PARITY_CASES[i] for i in (0, 1, 4, 5, 6, 8, 10)]


@pytest.mark.parametrize("name,kw", TASK_CASES, ids=[c[0] + "-" + "x".join(str(v) for v in c[1].values()) for c in TASK_CASES])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_boundary_entries_in_the_fused_kernel_are_bit_identical_to_the_list_kernels(name, kw, dtype, monkeypatch):
    """boundary entries evaluated by the fused kernel (task table, boundary.plan_tasks: ONE launch per
    step) against the list kernels (one launch per method + fused kernel): same arithmetic, so the
    interior populations must be IDENTICAL -- single steps, graph pairs, a stand-alone
    boundary_condition() in between, an outside write of F; time-dependent right-hand sides included
    (rayleigh_benard)."""
    import pylbm_b200
    from pylbm_b200 import cases, runtime as rt

    def run(tasks):
        monkeypatch.setenv("PYLBM_B200_TASKS", "1" if tasks else "0")
        monkeypatch.setenv("PYLBM_B200_NO_WALLS", "1")
        sim = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), dtype=dtype)
        assert (sim.bc.tasks is not None) == tasks
        before = rt.lib().lbm_sim_launch_count(sim._handle)
        sim.run(8)
        launches = rt.lib().lbm_sim_launch_count(sim._handle) - before
        sim.one_time_step()
        sim.boundary_condition()
        sim.F_halo[1] = sim.F_halo[1]
        sim.run(5)
        sim.one_time_step()
        return sim, launches

    (a, la), (b, lb) = run(True), run(False)
    assert a.bc.tasks["ntasks"] > 0
    if not a._time_dependent:
        assert la < lb            # first step: copy kernels for the periodic axes, then 1 launch per step
    inner = (slice(None),) + tuple(slice(v, -v) for v in a.domain.stencil.vmax)
    fluid = a.domain.in_or_out[inner[1:]] == a.domain.valin
    Fa, Fb = a.container.F.get()[inner], b.container.F.get()[inner]
    assert np.array_equal(Fa[:, fluid], Fb[:, fluid])
    for key in a.scheme.consm:
        assert np.array_equal(a.m[key][fluid], b.m[key][fluid])


def test_time_dependent_boundary_values_on_the_device(monkeypatch):
    """time_bc labels (reference: boundary.py:307-321): the device path (asynchronous upload of the
    callback's moments, equilibrium + m2f + rhs recomputation enqueued on the simulation's stream, no
    synchronisation) gives the populations of the host path (NumPy set_rhs + synchronous upload) bit for
    bit, with the task table and with the list kernels."""
    import pylbm_b200
    from pylbm_b200 import cases

    def run(host, tasks):
        if host:
            monkeypatch.setenv("PYLBM_B200_HOST_TIME_BC", "1")
        else:
            monkeypatch.delenv("PYLBM_B200_HOST_TIME_BC", raising=False)
        monkeypatch.setenv("PYLBM_B200_TASKS", "1" if tasks else "0")
        sim = pylbm_b200.Simulation(cases.rayleigh_benard(nx=64, ny=32, period=0.05, perturb=0))
        assert sim._time_dependent
        assert any(getattr(m, "_time_plans", None) for m in sim.bc.methods) == (not host)
        for _ in range(12):
            sim.one_time_step()
        sim.boundary_condition()
        sim.run(9)
        return sim.container.F.get()[:, 1:-1, 1:-1]

    ref = run(True, False)
    assert np.array_equal(run(False, False), ref)
    assert np.array_equal(run(False, True), ref)


def test_m_halo_setter_keeps_the_other_moments():
    """`sol.m_halo[k] = v` on the lazily allocated moment array: the other rows keep the moments of the
    last f2m, like the reference's host array (simulation.py:226-229)."""
    import pylbm_b200
    from pylbm_b200 import cases

    sim = pylbm_b200.Simulation(cases.karman_d2q9(nx=64, ny=32, perturb=0))
    sim.run(4)
    qx = sim.m_halo[cases.QX].copy()
    rho = sim.m_halo[cases.RHO].copy()
    sim.container.release_m()
    sim._update_m = True
    sim.m_halo[cases.RHO] = 2.0 * rho
    assert np.array_equal(sim.m_halo[cases.QX], qx)
    assert np.array_equal(sim.m_halo[cases.RHO], 2.0 * rho)


def test_a_symbol_without_a_value_fails_at_the_first_step():
    """a kernel scalar that is neither in 'parameters' nor in `extra_parameters` must raise, not run with
    zero (the reference fails in call_genfunction: algorithm/base.py:662-681)."""
    import sympy as sp
    import pylbm_b200
    from pylbm_b200 import cases

    grav = sp.Symbol("gravity_coeff")
    dico = cases.rayleigh_benard(nx=32, ny=16, time_bc=False)
    dico["schemes"][0]["source_terms"] = {cases.QY: grav * cases.T}
    sim = pylbm_b200.Simulation(dico)
    with pytest.raises(KeyError):
        sim.one_time_step()
    with pytest.raises(KeyError):
        sim.run(2)
    sim.extra_parameters[grav] = 0.01
    sim.one_time_step()
    assert sim.nt == 1
