"""
Host-side flow of `pylbm.Simulation(dico, generator='cuda')` (pylbm_b200/plugin.py) on a CPU box, with
the runtime replaced by a recording test double (tests/fake_runtime.py): the UNCHANGED reference
constructor must run to the end, hand the boundary lists of the reference to the runtime once, and a
time step must be exactly one runtime call.  The numbers are checked on the GPU (tests/test_gpu_plugin.py).
"""
import numpy as np
import pytest

import fake_runtime


@pytest.fixture
def cuda_pylbm(pylbm, monkeypatch):
    from pylbm_b200 import plugin

    fake = fake_runtime.install(monkeypatch)
    plugin.register()
    yield pylbm, fake
    # objects built on the test double must be finalised while it is still installed (their __del__ hands
    # pointers back to the runtime)
    import gc

    gc.collect()


@pytest.mark.parametrize("case,kw", [("karman_d2q9", dict(nx=32, ny=16)), ("lid_cavity_d3q19", dict(n=8)),
                                     ("rayleigh_benard", dict(nx=32, ny=16, period=0.2))])
@pytest.mark.parametrize("lowering", ["scheme", "ir"])
def test_reference_constructor_drives_the_cuda_backend(cuda_pylbm, case, kw, lowering):
    from pylbm_b200 import cases, plugin
    from pylbm_b200.simulation import CudaEngine

    pylbm, fake = cuda_pylbm
    dico = cases.CASES[case](mod=pylbm, generator="cuda", **kw)
    dico["cuda_option"] = {"lowering": lowering}
    sol = pylbm.Simulation(dico)
    assert isinstance(sol, pylbm.Simulation) and isinstance(sol, CudaEngine)
    assert type(sol).__name__ == "CudaSimulation"
    assert sol.generator.backend == "CUDA" and sol.generator.module.lowering == lowering
    # the reference's own front-end objects are in place
    assert type(sol.scheme).__module__.startswith("pylbm.") and type(sol.algo).__module__.startswith("pylbm.")
    assert type(sol.domain.geom).__module__.startswith("pylbm.")
    if lowering == "scheme":
        # the reference's symbolic routines are not consumed by this lowering: built on demand only
        assert not sol.generator.routines
        sol.algo.generate_routines()
    assert set(sol.generator.routines) >= {"one_time_step", "f2m", "m2f", "equilibrium", "relaxation", "transport"}
    for name in ("one_time_step", "f2m", "f2m_consm", "m2f", "equilibrium", "relaxation", "transport"):
        assert name in sol.kernels.info["routines"], name
    # boundary lists: same as the reference's Cython run builds from its dense Domain
    ref = pylbm.Simulation(cases.CASES[case](mod=pylbm, generator="numpy", **kw))
    assert len(sol.bc.methods) == len(ref.bc.methods)
    for a, b in zip(sol.bc.methods, ref.bc.methods):
        assert type(a).__name__ == type(b).__name__
        order = np.argsort(a._order) if a._order is not None else slice(None)
        assert np.array_equal(a.istore, b.istore) and np.array_equal(a.ilabel, b.ilabel)
        for la, lb in zip(a.iload, b.iload):
            assert np.array_equal(la, lb)
        assert np.allclose(a.distance, b.distance, rtol=0, atol=0)
    assert np.array_equal(sol.domain.in_or_out, ref.domain.in_or_out)
    # one runtime call per step, nothing else
    nbc = fake.count("lbm_sim_add_bc")
    assert nbc >= len(sol.bc.methods)
    before = len(fake.calls)
    time_dependent = any(m.is_time_dependent for m in sol.bc.methods)
    for _ in range(3):
        sol.one_time_step()
    new = [n for n, _ in fake.calls[before:]]
    assert new.count("lbm_sim_step") == 3
    if not time_dependent:
        assert new == ["lbm_sim_step"] * 3, new
    assert sol.nt == 3 and abs(sol.t - 3 * sol.dt) < 1e-15
    sol.boundary_condition()
    assert fake.calls[-1][0] == "lbm_sim_boundary_condition"


def test_other_generators_are_untouched(cuda_pylbm):
    from pylbm_b200 import cases
    from pylbm_b200.simulation import CudaEngine

    pylbm, fake = cuda_pylbm
    sol = pylbm.Simulation(cases.karman_d2q9(nx=32, ny=16, mod=pylbm, generator="numpy"))
    assert type(sol) is pylbm.Simulation and not isinstance(sol, CudaEngine)
    assert type(sol.domain).__module__ == "pylbm.domain" and type(sol.bc).__module__ == "pylbm.boundary"
    sol.one_time_step()
    assert not any(n.startswith("lbm_sim") for n, _ in fake.calls)


def test_no_gpu_is_an_error_not_a_fallback(pylbm):
    from pylbm_b200 import cases, plugin, runtime

    if runtime.lib().lbm_device_count() > 0:
        pytest.skip("this box has a GPU")
    plugin.register()
    with pytest.raises(runtime.LbmError):
        pylbm.Simulation(cases.karman_d2q9(nx=32, ny=16, mod=pylbm, generator="cuda"))


def test_in_place_option_registers_odd_lists_and_one_array(cuda_pylbm):
    """`dico['cuda_option'] = {'in_place': True}`: ONE population array, the odd-step lists of every method
    and the even / odd launcher are handed to the runtime; a step is still one runtime call."""
    from pylbm_b200 import cases

    pylbm, fake = cuda_pylbm
    dico = cases.karman_d2q9(nx=32, ny=16, mod=pylbm, generator="cuda")
    dico["cuda_option"] = {"in_place": True}
    sol = pylbm.Simulation(dico)
    assert sol.container.Fnew is sol.container.F and sol.in_place
    assert "one_time_step_aa" not in sol.kernels.info["routines"]          # a launcher, not a routine
    assert {"f2m_sw", "f2m_consm_sw"} <= set(sol.kernels.info["routines"])
    assert fake.count("lbm_sim_set_bc_odd") == len(sol.bc.methods) and fake.count("lbm_sim_set_aa") == 1
    assert fake.count("lbm_sim_set_walls") == 0 and fake.count("lbm_sim_set_tasks") == 0
    before = len(fake.calls)
    sol.one_time_step()
    assert [n for n, _ in fake.calls[before:]] == ["lbm_sim_step"]
    # the odd lists are the even lists moved by (k, y) -> (kbar, y + v_k): same length, all different
    for method, (store, loads) in zip(sol.bc.methods, sol._odd_keep):
        assert store.shape == method._keep[0].shape and (store != method._keep[0]).all()
