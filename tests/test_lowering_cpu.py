"""
Host logic of the CUDA backend, checked without a GPU: the optimised per-cell IR
(algorithm.py: T(u)(M f) factorisation) after SSA + CSE (cudagen.lower_statements) evaluated with
NumPy must agree with the oracle's literal restatement of the reference kernel (dense matrices,
sequential statements) compiled to C.
"""
import ctypes

import numpy as np
import pytest

from lowering_eval import evaluate

CASES = [
    ("lid_cavity_d2q9", dict(n=16)),
    ("karman_d2q9", dict(nx=32, ny=16, relative_velocity=False)),
    ("shallow_water_d2q4", dict(n=16)),
    ("lid_cavity_d3q19", dict(n=8)),
    ("channel_sphere_d3q27", dict(nx=16, ny=8, nz=8)),
]


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_fused_kernel_ir_matches_literal_restatement(name, kw):
    from pylbm_b200 import cases, cudagen
    from pylbm_b200.algorithm import PullAlgorithm
    from pylbm_b200.scheme import Scheme
    from oracle.lbm_oracle import build_library

    scheme = Scheme(cases.CASES[name](**kw))
    algo = PullAlgorithm(scheme)
    ir = algo.one_time_step()
    lib, _ = build_library(scheme)
    dim, Q = scheme.dim, algo.ns
    n = [6] * dim + [1] * (3 - dim)
    vmax = list(scheme.stencil.vmax) + [0] * (3 - dim)
    rng = np.random.default_rng(1)
    f = 1.0 / Q + 0.01 * rng.uniform(-1, 1, size=tuple(n) + (Q,))
    fnew = np.zeros_like(f)
    lib.one_time_step(f.ctypes.data_as(ctypes.c_void_p), fnew.ctypes.data_as(ctypes.c_void_p),
                      *[ctypes.c_int(v) for v in n], ctypes.c_double(0.0), ctypes.c_double(0.1),
                      (ctypes.c_double * 1)(0.0))
    inner = tuple(slice(v, nn - v) for v, nn in zip(vmax, n))
    pulled = []
    for k in range(Q):
        off = list(-algo.velocities[k]) + [0] * (3 - dim)
        sl = tuple(slice(v + o, nn - v + o) for v, nn, o in zip(vmax, n, off))
        pulled.append(f[sl + (k,)])
    out = evaluate(ir, pulled, {"dt": 0.1, "t": 0.0})
    err = max(np.abs(out[k] - fnew[inner + (k,)]).max() for k in range(Q))
    assert err <= 5e-15
    # and without CSE (plain SSA) the result is the same to rounding
    out2 = evaluate(ir, pulled, {"dt": 0.1, "t": 0.0}, cse=False)
    assert max(np.abs(a - b).max() for a, b in zip(out, out2)) <= 5e-15
    add, mul, div = cudagen.count_ops(*cudagen.lower_statements(ir.statements, ir.outputs))
    assert add + mul < 60 * Q   # the lowering keeps the kernel far below the reference's ~66 Q ops/cell


def test_generated_source_is_deterministic_and_describes_itself():
    from pylbm_b200 import cases, cudagen
    from pylbm_b200.algorithm import PullAlgorithm
    from pylbm_b200.scheme import Scheme

    scheme = Scheme(cases.lid_cavity_d3q19(n=16))
    a, ia = cudagen.generate_source(PullAlgorithm(scheme).kernels(), 3, 19)
    b, ib = cudagen.generate_source(PullAlgorithm(Scheme(cases.lid_cavity_d3q19(n=16))).kernels(), 3, 19)
    assert a == b and ia["hash"] == ib["hash"]
    assert set(ia["routines"]) == {"transport", "f2m", "f2m_consm", "m2f", "relaxation", "equilibrium", "one_time_step"}
    assert "lbmk_kernel_one_time_step" in a and "__launch_bounds__" in a
