import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available():
    try:
        from pylbm_b200 import runtime

        return runtime.lib().lbm_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """tests marked `gpu` are skipped (not failed) on a box without a CUDA device or without nvcc."""
    gpu_items = [item for item in items if "gpu" in item.keywords]
    if not gpu_items or _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (or the runtime library cannot be built) on this box")
    for item in gpu_items:
        item.add_marker(skip)


def reference_paths():
    """sys.path entries that make the UNMODIFIED reference importable, or None: oracle/_ref (installed
    by tools/make_ref.sh, travels to the GPU box with the working tree), else the read-only checkout of
    the build container with the import shims of tools/refshim."""
    installed = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(os.path.join(installed, "pylbm")):
        return [os.path.join(installed, "shims"), installed]
    checkout = os.environ.get("PYLBM_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(checkout, "pylbm")):
        return [os.path.join(ROOT, "tools", "refshim"), checkout]
    return None


@pytest.fixture(scope="session")
def pylbm():
    """the reference package (skips when it is not on this box)."""
    paths = reference_paths()
    if paths is None:
        pytest.skip("reference pylbm not available on this box (run tools/make_ref.sh in the build container)")
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pylbm as ref

    return ref


def case_id(name, kw):
    return name + "-" + "x".join(str(v) for v in kw.values())


@pytest.fixture(scope="session")
def small_cases():
    """(name, kwargs) of the parity workloads at sizes the oracle finishes in seconds."""
    return PARITY_CASES


PARITY_CASES = [
    ("lid_cavity_d2q9", dict(n=32)),
    ("karman_d2q9", dict(nx=128, ny=32)),
    ("karman_d2q9", dict(nx=96, ny=32, relative_velocity=False)),
    ("shallow_water_d2q4", dict(n=32)),
    ("lid_cavity_d3q19", dict(n=16)),
    ("channel_sphere_d3q27", dict(nx=32, ny=16, nz=16)),
    # features of the reference demos not covered by the BASELINE configs
    ("rayleigh_benard", dict(nx=64, ny=32, period=0.2)),   # 2 schemes, source term, Bouzidi anti-BB, time_bc
    ("advection_d1q5", dict(n=64)),                        # 1-D, ghost width 2, Neumann, bounce-back value
    ("heat_d2q5", dict(n=32)),                             # anti-BB, Bouzidi anti-BB, NeumannY, triangle, ellipse
    ("advection_d3q6", dict(n=12)),                        # 3-D periodic, init on distributions
    # edge cases: ragged sizes (not multiples of the block / alignment), two ghost layers in 2-D
    ("advection_d2q13", dict(nx=23, ny=19)),               # D2Q13: vmax = 2 on both axes, BB value + NeumannX
    ("channel_sphere_d3q27", dict(nx=21, ny=13, nz=9)),    # 3-D, every extent odd
]
