"""
The north-star interface on the GPU: the UNCHANGED `pylbm.Simulation(dico, generator='cuda')` of an
importable reference (oracle/_ref, installed by tools/make_ref.sh), after `pylbm_b200.plugin.register()`.

* against the fixtures of the unmodified reference's Cython generator (tests/golden/ref_*.npz: boundary
  lists `array_equal`, rhs <= 1e-15, conserved moments after 50 steps <= 1e-12 of max|field|) on the 12
  parity workloads, both lowerings;
* against ALL the reference's demo regression fixtures (tests/golden/demos: the 26 demo tests, 3 more demos,
  7 notebook simulations -- Bouzidi, time-dependent boundary values, vectorial schemes, D3Q6/D3Q15, source
  terms) with the demos' own loop;
* against the reference's Cython generator run LIVE in the same process on the same dictionary;
* the kwargs protocol (`sol.algo.call_function`) and the item properties keep the reference's meaning.
fp64 tolerance 1e-12 (BASELINE.json north_star).
"""
import os

import numpy as np
import pytest

from conftest import PARITY_CASES, case_id
from demo_fixtures import final_fields, load_demo, load_results, run_to_final_time

pytestmark = pytest.mark.gpu
TOL = 1e-12
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IDS = [case_id(*c) for c in PARITY_CASES]


@pytest.fixture(scope="module")
def cuda(pylbm):
    from pylbm_b200 import plugin

    plugin.register()
    return pylbm


def _check_against_fixture(sol, ref, steps=None):
    assert len(sol.bc.methods) == int(ref["nmethods"])
    for i, method in enumerate(sol.bc.methods):
        pre = "bc%d_" % i
        assert type(method).__name__ == str(ref[pre + "name"])
        assert method.istore.dtype == np.int32 and np.array_equal(method.istore, ref[pre + "istore"])
        for j, il in enumerate(method.iload):
            assert np.array_equal(il, ref[pre + "iload%d" % j])
        if hasattr(method, "s"):
            assert np.array_equal(method.s, ref[pre + "s"])
        np.testing.assert_allclose(method.rhs, ref[pre + "rhs"], rtol=0, atol=1e-15)
    for _ in range(int(ref["nsteps"]) if steps is None else steps):
        sol.one_time_step()                      # the call a pylbm user makes
    fluid = ref["in_or_out"][tuple(slice(v, -v) for v in sol.domain.stencil.vmax)] == sol.domain.valin
    worst = 0.0
    for key in sol.scheme.consm:
        a, b = sol.m[key], ref["m_" + str(key)]
        worst = max(worst, np.abs(a[fluid] - b[fluid]).max() / np.abs(b[fluid]).max())
    return worst


@pytest.mark.parametrize("name,kw", PARITY_CASES, ids=IDS)
def test_pylbm_simulation_cuda_against_reference_fixture(cuda, name, kw):
    from pylbm_b200 import cases
    from pylbm_b200.simulation import CudaEngine

    ref = np.load(os.path.join(GOLDEN, "ref_%s.npz" % case_id(name, kw)))
    sol = cuda.Simulation(cases.CASES[name](perturb=0, mod=cuda, generator="cuda", **kw))
    assert isinstance(sol, cuda.Simulation) and isinstance(sol, CudaEngine)
    assert type(sol.scheme).__module__ == "pylbm.scheme"          # the reference's own front end
    worst = _check_against_fixture(sol, ref)
    assert worst <= TOL, worst
    assert sol.nt == int(ref["nsteps"])


@pytest.mark.parametrize("name,kw", [PARITY_CASES[i] for i in (1, 3, 4, 5, 6, 10)], ids=[IDS[i] for i in (1, 3, 4, 5, 6, 10)])
def test_reference_ir_lowering_against_reference_fixture(cuda, name, kw):
    """the kernels lowered from the reference's OWN symbolic routines (cuda_option lowering='ir')."""
    from pylbm_b200 import cases

    ref = np.load(os.path.join(GOLDEN, "ref_%s.npz" % case_id(name, kw)))
    dico = cases.CASES[name](perturb=0, mod=cuda, generator="cuda", **kw)
    dico["cuda_option"] = {"lowering": "ir"}
    sol = cuda.Simulation(dico)
    assert sol.generator.module.lowering == "ir"
    worst = _check_against_fixture(sol, ref)
    assert worst <= TOL, worst


# every dictionary of the reference's demo regression suite and tutorial notebooks (tests/golden/demos)
from demo_fixtures import demo_names

DEMOS = demo_names()


@pytest.mark.parametrize("test", DEMOS)
def test_pylbm_simulation_cuda_reproduces_reference_demo(cuda, test):
    dico, kwargs, record = load_demo(test, mod=cuda, generator="cuda")
    expected = load_results(test)
    sol = cuda.Simulation(dico, **kwargs)
    run_to_final_time(sol, record["final_time"])
    assert sol.nt == expected["nsteps"]
    got = final_fields(sol, expected["plane_stride"])
    tol = TOL * max(1.0, expected["nsteps"] / 50.0)
    for kind in ("ref", "h5"):
        fields = expected[kind]
        if fields is None:
            continue
        for key, want in fields.items():
            err = np.abs(got[key] - want).max() / max(np.abs(want).max(), 1e-300)
            assert err <= tol, (test, kind, key, err)


@pytest.mark.parametrize("name,kw", [("karman_d2q9", dict(nx=96, ny=32)), ("lid_cavity_d3q19", dict(n=12))])
def test_cuda_against_the_reference_cython_generator_live(cuda, name, kw, tmp_path):
    """same process, same dictionary (seeded perturbed initial state), generator='cuda' vs the reference's
    own generator='cython': lists equal, conserved moments within 1e-12 after 40 steps."""
    from pylbm_b200 import cases

    d_ref = cases.CASES[name](perturb=7, mod=cuda, generator="cython", **kw)
    d_ref["codegen_option"] = {"directory": str(tmp_path)}
    ref = cuda.Simulation(d_ref)
    sol = cuda.Simulation(cases.CASES[name](perturb=7, mod=cuda, generator="cuda", **kw))
    assert type(ref) is cuda.Simulation and type(sol) is not cuda.Simulation
    for a, b in zip(sol.bc.methods, ref.bc.methods):
        assert np.array_equal(a.istore, b.istore) and np.array_equal(a.iload[0], b.iload[0])
        np.testing.assert_allclose(a.rhs, b.rhs, rtol=0, atol=1e-15)
    for _ in range(40):
        sol.one_time_step()
        ref.one_time_step()
    fluid = ref.domain.in_or_out[tuple(slice(v, -v) for v in ref.domain.stencil.vmax)] == ref.domain.valin
    for key in ref.scheme.consm:
        a, b = sol.m[key], ref.m[key]
        err = np.abs(a[fluid] - b[fluid]).max() / np.abs(b[fluid]).max()
        assert err <= TOL, (str(key), err)


def test_reference_protocols_on_the_cuda_simulation(cuda):
    """what reference user code touches besides one_time_step: `sol.algo.call_function` (the kwargs
    protocol of symbolic.py:288-299 on the module object), the split step of simulation.py:392-408,
    item properties, `run(n)`."""
    import sympy as sp
    from pylbm_b200 import cases

    kw = dict(n=12)
    a = cuda.Simulation(cases.lid_cavity_d3q19(perturb=1, mod=cuda, generator="cuda", **kw))
    b = cuda.Simulation(cases.lid_cavity_d3q19(perturb=1, mod=cuda, generator="cuda", **kw))
    c = cuda.Simulation(cases.lid_cavity_d3q19(perturb=1, mod=cuda, generator="cuda", **kw))
    for _ in range(3):
        a.one_time_step()
        b.boundary_condition()
        for name in ("transport", "f2m", "relaxation", "m2f"):
            getattr(b, name)()
    c.run(3)
    inner = (slice(None),) + (slice(1, -1),) * 3
    Fa, Fb, Fc = (s.container.F.get()[inner] for s in (a, b, c))
    assert np.abs(Fa - Fb).max() <= 1e-13 * np.abs(Fa).max()
    assert np.array_equal(Fa, Fc) and c.nt == 3
    # the reference's own dispatcher on the module object: f2m through call_function == sol.f2m()
    a.f2m()
    m_engine = a.container.m.get()
    a.container.m.set(np.zeros_like(m_engine))
    a.algo.call_function("f2m", a)
    assert np.array_equal(a.container.m.get(), m_engine)
    mass = list(a.scheme.consm)[0]            # the density (first conserved moment)
    assert a.m[mass].shape == tuple(a.domain.shape_in) and a.m_halo[mass].shape == tuple(a.domain.shape_halo)
    assert a.F[0].shape == tuple(a.domain.shape_in)
    assert abs(a.m[mass].mean() - 1.0) < 1e-3
    assert isinstance(str(a), str) and "Simulation" in str(a)


def test_fp32_storage_through_the_reference_dtype_argument(cuda):
    """`pylbm.Simulation(dico, dtype='float32')`: the reference accepts and ignores dtype
    (simulation.py:89-91); the CUDA backend stores the populations in fp32 (tolerance 5e-5, 50 steps)."""
    from pylbm_b200 import cases

    name, kw = PARITY_CASES[4]
    ref = np.load(os.path.join(GOLDEN, "ref_%s.npz" % case_id(name, kw)))
    sol = cuda.Simulation(cases.CASES[name](perturb=0, mod=cuda, generator="cuda", **kw), dtype="float32")
    assert sol.container.F.storage == "f32"
    worst = _check_against_fixture(sol, ref)
    assert worst <= 5e-5, worst
