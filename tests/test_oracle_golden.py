"""
Pins the oracle (and the host front-end it shares with the product) against the REFERENCE:
fixtures of tests/golden/ were produced by tools/make_golden.py from the unmodified pylbm 0.11.0
(Cython generator) and from the reference's own golden HDF5 fields.  CPU only.
"""
import json
import os

import numpy as np
import pytest

from conftest import PARITY_CASES, case_id

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


def _oracle(case, kw, **more):
    from pylbm_b200 import cases
    from oracle.lbm_oracle import OracleSimulation

    return OracleSimulation(cases.CASES[case](generator="cuda", **kw, **more))


def _rel_err(a, b, mask):
    scale = max(np.abs(b[mask]).max(), 1e-300)
    return np.abs(a[mask] - b[mask]).max() / scale


@pytest.mark.parametrize("case,kw", PARITY_CASES, ids=[case_id(*c) for c in PARITY_CASES])
def test_reference_fixture(case, kw):
    ref = np.load(os.path.join(GOLDEN, "ref_%s.npz" % case_id(case, kw)))
    sim = _oracle(case, kw, perturb=0)

    # ---- host front-end: bit-exact against the reference ----
    assert np.array_equal(sim.domain.in_or_out, ref["in_or_out"])
    assert len(sim.bc.methods) == int(ref["nmethods"])
    for i, method in enumerate(sim.bc.methods):
        pre = "bc%d_" % i
        assert type(method).__name__ == str(ref[pre + "name"])
        assert method.istore.dtype == ref[pre + "istore"].dtype == np.int32
        assert np.array_equal(method.istore, ref[pre + "istore"])
        for j, il in enumerate(method.iload):
            assert np.array_equal(il, ref[pre + "iload%d" % j])
        assert np.array_equal(method.ilabel, ref[pre + "ilabel"])
        assert np.array_equal(method.distance, ref[pre + "distance"])
        if hasattr(method, "s"):
            assert np.array_equal(method.s, ref[pre + "s"])
        # rhs goes through the generated equilibrium / m2f arithmetic: 1 ulp
        np.testing.assert_allclose(method.rhs, ref[pre + "rhs"], rtol=0, atol=2e-16 * max(1.0, np.abs(ref[pre + "rhs"]).max()) * 4)
    params = list(sim.scheme.param.items())
    M = np.array(sim.scheme.M.subs(params).tolist(), dtype=float)
    invM = np.array(sim.scheme.invM.subs(params).tolist(), dtype=float)
    np.testing.assert_allclose(M, ref["M"], rtol=1e-15, atol=1e-15)
    np.testing.assert_allclose(invM, ref["invM"], rtol=1e-14, atol=1e-15)

    # ---- initial state and NSTEPS steps ----
    inner = tuple(slice(v, -v) for v in sim.domain.stencil.vmax)
    fluid = sim.domain.in_or_out[inner] == sim.domain.valin
    for key in sim.scheme.consm:
        assert _rel_err(sim.m[key], ref["m0_" + str(key)], fluid) <= 1e-13
    for _ in range(int(ref["nsteps"])):
        sim.one_time_step()
    for key in sim.scheme.consm:
        err = _rel_err(sim.m[key], ref["m_" + str(key)], fluid)
        assert err <= TOL, (str(key), err)


def test_reference_golden_h5_fields():
    """the reference's own golden files (dx = 1/64, Tf = 0.5, solid cells zeroed;
    tolerance of the reference's h5diff: atol 1e-7, rtol 1e-14 -- we ask for 1e-12 absolute)."""
    manifest = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))
    checked = 0
    for name, meta in manifest["files"].items():
        if not name.startswith("h5_"):
            continue
        ref = np.load(os.path.join(GOLDEN, name))
        sim = _oracle(meta["case"], meta["kwargs"])
        assert abs(sim.domain.dx - meta["space_step"]) < 1e-15
        while sim.t < meta["final_time"]:
            sim.one_time_step()
        inner = tuple(slice(v, -v) for v in sim.domain.stencil.vmax)
        solid = sim.domain.in_or_out[inner] != sim.domain.valin
        for key in sim.scheme.consm:
            field = sim.m[key].copy()
            field[solid] = 0
            assert np.abs(field - ref[str(key)]).max() <= 1e-12, (name, str(key))
            checked += 1
        for d in range(sim.dim):
            np.testing.assert_allclose(sim.domain.coords[d], ref["x_%d" % d], rtol=0, atol=1e-14)
    assert checked >= 9
