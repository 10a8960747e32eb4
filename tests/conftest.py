import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
]
