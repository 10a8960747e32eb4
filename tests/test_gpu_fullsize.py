"""
BASELINE.json's full-size configurations, checked through size-independent properties (the oracle
cannot run them in seconds): conservation, symmetry, determinism of the enqueue paths.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_gb():
    import ctypes
    from pylbm_b200 import runtime as rt

    free, total = ctypes.c_uint64(), ctypes.c_uint64()
    rt.check(rt.lib().lbm_mem_info(ctypes.byref(free), ctypes.byref(total)), "lbm_mem_info")
    return free.value / 1e9


def test_c4_d3q19_512_mass_and_mirror_symmetry():
    """D3Q19 MRT lid-driven cavity 512^3 (config 4): bounce-back walls conserve the total mass (the
    lid only adds tangential momentum), and the flow driven along x stays mirror-symmetric in y."""
    import pylbm_b200
    from pylbm_b200 import cases

    if _free_gb() < 80:
        pytest.skip("needs ~65 GB of HBM")
    sim = pylbm_b200.Simulation(cases.lid_cavity_d3q19(n=512))
    assert sim.domain.shape_in == [512, 512, 512]
    assert sum(m.istore.shape[0] for m in sim.bc.methods) > 6 * 510 * 510 * 5
    mass0 = float(sim.m[cases.RHO].sum(dtype=np.float64))
    sim.run(40)
    rho = sim.m[cases.RHO]
    assert abs(float(rho.sum(dtype=np.float64)) - mass0) <= 1e-9 * mass0
    qy = sim.m[cases.QY]
    qx = sim.m[cases.QX]
    assert np.abs(qx).max() > 1e-4                       # the lid has started to drag the fluid
    scale = np.abs(qx).max()
    assert np.abs(qy + qy[:, ::-1, :]).max() <= 1e-12 * scale
    assert np.abs(qx - qx[:, ::-1, :]).max() <= 1e-12 * scale


def test_c3_shallow_water_4096_conservation_and_graph_path():
    """D2Q4x3 shallow water 4096^2 (config 3), fully periodic: every conserved moment is conserved;
    run(n) (CUDA-graph pairs) is bit-identical to n single steps."""
    import pylbm_b200
    from pylbm_b200 import cases

    a = pylbm_b200.Simulation(cases.shallow_water_d2q4(n=4096))
    b = pylbm_b200.Simulation(cases.shallow_water_d2q4(n=4096))
    before = {k: float(a.m[k].sum(dtype=np.float64)) for k in a.scheme.consm}
    a.run(31)
    for _ in range(31):
        b.one_time_step()
    for k, v in before.items():
        fa, fb = a.m[k], b.m[k]
        assert np.array_equal(fa, fb)
        assert abs(float(fa.sum(dtype=np.float64)) - v) <= 1e-9 * max(1.0, abs(v))


def test_c2_karman_4096x1024_boundary_idempotent():
    """D2Q9 Karman 4096x1024 (config 2): the boundary step only depends on interior values, so
    applying it twice changes nothing; the obstacle list is symmetric in size about its centre line."""
    import pylbm_b200
    from pylbm_b200 import cases

    sim = pylbm_b200.Simulation(cases.karman_d2q9(nx=4096, ny=1024))
    sim.run(10)
    sim.boundary_condition()
    once = sim.container.F.get()
    sim.boundary_condition()
    twice = sim.container.F.get()
    assert np.array_equal(once, twice)
    labels = np.concatenate([m.ilabel for m in sim.bc.methods])
    assert (labels == 2).sum() > 1000 and (labels == 1).sum() == 3 * 1024


def test_more_than_65535_rows():
    """maximum sizes: a long 2-D lattice has more row groups than a CUDA grid allows along y/z; the
    launcher folds (row group, plane) into one linear index.  D2Q9 Karman 70000 x 130 against the
    oracle after 3 steps (boundary lists on all four faces and the obstacle included)."""
    import pylbm_b200
    from pylbm_b200 import cases
    from oracle.lbm_oracle import OracleSimulation

    kw = dict(nx=70000, ny=130)
    sim = pylbm_b200.Simulation(cases.karman_d2q9(perturb=cases.WAVE, **kw))
    ora = OracleSimulation(cases.karman_d2q9(perturb=cases.WAVE, **kw))
    sim.run(3)
    for _ in range(3):
        ora.one_time_step()
    fluid = ora.domain.in_or_out[1:-1, 1:-1] == ora.domain.valin
    for key in sim.scheme.consm:
        a, b = sim.m[key], ora.m[key]
        err = np.abs(a[fluid] - b[fluid]).max() / np.abs(b[fluid]).max()
        assert err <= 1e-12, (str(key), err)


@pytest.mark.parametrize("name,kw,steps", [
    ("lid_cavity_d3q19", dict(n=128), 24),
    ("channel_sphere_d3q27", dict(nx=128, ny=64, nz=64), 16),
])
def test_mid_size_3d_against_the_oracle_with_fused_walls(name, kw, steps):
    """the sizes where the production configuration of the 3-D workloads is active -- walls of the
    fastest axis applied by the fused kernel (boundary.plan_walls), merged list launches, CUDA-graph
    pairs -- compared cell by cell with the oracle (OpenMP build of the same restatement): conserved
    moments on fluid cells within 1e-12 of max|field| (seeded perturbed initial state)."""
    import pylbm_b200
    from pylbm_b200 import cases
    from oracle.lbm_oracle import OracleSimulation

    sim = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw))
    assert sim.bc.walls is not None, "the fused-wall plan must be active for this configuration"
    ora = OracleSimulation(cases.CASES[name](perturb=0, **kw), openmp=True)
    sim.run(steps)
    for _ in range(steps):
        ora.one_time_step()
    fluid = ora.domain.in_or_out[1:-1, 1:-1, 1:-1] == ora.domain.valin
    for key in sim.scheme.consm:
        okey = [k for k in ora.scheme.consm if str(k) == str(key)][0]
        a, b = sim.m[key], ora.m[okey]
        err = np.abs(a[fluid] - b[fluid]).max() / np.abs(b[fluid]).max()
        assert err <= 1e-12, (name, str(key), err)
