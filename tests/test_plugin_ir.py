"""
`generator='cuda'` registered inside the reference pylbm (pylbm_b200/plugin.py).

Needs the reference importable (oracle/_ref installed by tools/make_ref.sh, else /root/reference
through tools/refshim: the `pylbm` fixture of conftest.py); skipped elsewhere.  CPU only: the reference's own symbolic Routines are lowered to the per-cell IR, evaluated
with NumPy and compared with the literal C restatement; then `pylbm.Simulation(generator='cuda')` is
driven up to its first device allocation.
"""
import ctypes
import os
import sys

import numpy as np
import pytest


def _routines(pylbm, dico):
    """the Routines the reference hands to a backend, without compiling anything."""
    from pylbm.algorithm import PullAlgorithm
    from pylbm.generator import Generator

    scheme = pylbm.Scheme(dico)
    gen = Generator("CYTHON")
    sorder = list(range(scheme.dim + 1))
    PullAlgorithm(scheme, sorder, gen, {"m_local": True, "split": False, "check_isfluid": False}).generate()
    return scheme, gen.routines


@pytest.mark.parametrize("case,kw", [("karman_d2q9", dict(nx=32, ny=16)), ("lid_cavity_d3q19", dict(n=8)),
                                     ("rayleigh_benard", dict(nx=32, ny=16))])
def test_reference_routines_lower_to_the_same_kernel(pylbm, case, kw):
    from lowering_eval import evaluate
    from pylbm_b200 import cases, cudagen
    from pylbm_b200.plugin import routine_to_ir
    from pylbm_b200.scheme import Scheme
    from oracle.lbm_oracle import build_library

    ref_scheme, routines = _routines(pylbm, cases.CASES[case](mod=pylbm, generator="cython", **kw))
    assert {"transport", "f2m", "m2f", "relaxation", "equilibrium", "one_time_step"} <= set(routines)
    irs = {name: routine_to_ir(r) for name, r in routines.items()}
    ir = irs["one_time_step"]
    assert ir.in_array == "f" and ir.out_array == "fnew" and ir.inner
    vel = ref_scheme.stencil.get_all_velocities()
    assert [tuple(o) for o in ir.in_offsets] == [tuple(-int(c) for c in v) for v in vel]
    assert irs["f2m"].in_array == "f" and irs["f2m"].out_array == "m" and not irs["f2m"].inner
    assert irs["equilibrium"].in_array == irs["equilibrium"].out_array == "m"

    # numerics of the lowered reference IR == literal C restatement built from OUR scheme object
    scheme = Scheme(cases.CASES[case](**kw))
    lib, _ = build_library(scheme)
    dim, Q = scheme.dim, len(ir.in_syms)
    n = [6] * dim + [1] * (3 - dim)
    vmax = list(scheme.stencil.vmax) + [0] * (3 - dim)
    rng = np.random.default_rng(2)
    f = 1.0 / Q + 0.01 * rng.uniform(-1, 1, size=tuple(n) + (Q,))
    fnew = np.zeros_like(f)
    lib.one_time_step(f.ctypes.data_as(ctypes.c_void_p), fnew.ctypes.data_as(ctypes.c_void_p),
                      *[ctypes.c_int(v) for v in n], ctypes.c_double(0.0), ctypes.c_double(0.05),
                      (ctypes.c_double * 1)(0.0))
    inner = tuple(slice(v, nn - v) for v, nn in zip(vmax, n))
    pulled = []
    for k in range(Q):
        off = list(ir.in_offsets[k]) + [0] * (3 - dim)
        pulled.append(f[tuple(slice(v + o, nn - v + o) for v, nn, o in zip(vmax, n, off)) + (k,)])
    out = evaluate(ir, pulled, {"dt": 0.05, "t": 0.0})
    err = max(np.abs(out[k] - fnew[inner + (k,)]).max() for k in range(Q))
    assert err <= 5e-15
    # the whole set compiles to one CUDA translation unit
    source, info = cudagen.generate_source(list(irs.values()), dim, Q)
    assert set(info["routines"]) == set(irs)


def test_registered_backend_reaches_the_device(pylbm):
    """unchanged pylbm.Simulation + generator='cuda': on a box without GPU the construction must stop
    at the device (no silent fallback); with a GPU it must run and agree with the Cython generator."""
    from pylbm_b200 import cases, plugin, runtime

    plugin.register()
    dico = cases.karman_d2q9(nx=32, ny=16, mod=pylbm, generator="cuda")
    if runtime.lib().lbm_device_count() <= 0:
        with pytest.raises(runtime.LbmError):
            pylbm.Simulation(dico)
        return
    sol = pylbm.Simulation(dico)
    ref = pylbm.Simulation(cases.karman_d2q9(nx=32, ny=16, mod=pylbm, generator="cython"))
    for _ in range(10):
        sol.one_time_step()
        ref.one_time_step()
    for key in ref.scheme.consm:
        assert np.abs(sol.m[key] - ref.m[key]).max() <= 1e-12


def test_module_object_built_from_a_reference_simulation(pylbm):
    """every routine the reference registers for a full simulation (kernels + boundary loops) gets a
    callable with the kwargs protocol of symbolic.py:288-299; the CUDA source compiles with nvcc."""
    from pylbm_b200 import cases
    from pylbm_b200.plugin import CudaModule, BC_ROUTINES

    ref = pylbm.Simulation(cases.karman_d2q9(nx=32, ny=16, mod=pylbm, generator="cython"))
    routines = list(ref.generator.routines.values())
    module = CudaModule(routines)
    assert module.lowering == "ir"
    for r in routines:
        fn = getattr(module, r.name)
        assert callable(fn) and isinstance(fn.arg_dict, dict)
        wanted = {str(a.name) for a in r.arguments}
        if r.name in BC_ROUTINES:
            assert set(fn.arg_dict) <= wanted
        else:
            assert set(fn.arg_dict) <= wanted | {"dt", "t"}
    assert "lbmk_kernel_one_time_step" in module.source


DEMO_SCHEMES = ["test1D_euler", "test1D_advection_reaction", "test2D_orszag_Tang_vortex", "test2D_shallow_water",
                "test2D_kelvin_Helmoltz", "test2D_air_conditioning", "test3D_poseuille", "test3D_karman", "nb04_1"]


@pytest.mark.parametrize("test", DEMO_SCHEMES)
def test_reference_routines_of_the_demo_dictionaries(pylbm, test):
    """the plugin path on the dictionaries of the reference's own demos / notebooks: the Routines the
    REFERENCE's symbolic algorithm produces for them (vectorial schemes with up to six sub-schemes,
    relative velocities, source terms, D3Q15) are lowered by plugin.routine_to_ir and evaluated with
    NumPy; the result must equal the literal C restatement of the same scheme."""
    from demo_fixtures import load_demo
    from lowering_eval import evaluate
    from pylbm_b200.plugin import routine_to_ir
    from pylbm_b200.scheme import Scheme
    from oracle.lbm_oracle import build_library

    ref_dico, _, _ = load_demo(test, mod=pylbm, generator="cython")
    ref_scheme, routines = _routines(pylbm, ref_dico)
    ir = routine_to_ir(routines["one_time_step"])
    our_dico, _, _ = load_demo(test)
    scheme = Scheme(our_dico)
    lib, extras = build_library(scheme)
    assert not extras
    dim, Q = scheme.dim, len(ir.in_syms)
    assert Q == int(scheme.stencil.nv_ptr[-1])
    vel = scheme.stencil.get_all_velocities()
    assert [tuple(o) for o in ir.in_offsets] == [tuple(-int(c) for c in v) for v in vel]
    vmax = list(scheme.stencil.vmax) + [0] * (3 - dim)
    n = [4 + 2 * v for v in vmax[:dim]] + [1] * (3 - dim)
    rng = np.random.default_rng(5)
    f = 1.0 / Q + 0.01 * rng.uniform(-1, 1, size=tuple(n) + (Q,))
    # vectorial schemes divide by their own density (h, rho): keep every sub-scheme's mass near one
    nv_ptr = list(scheme.stencil.nv_ptr)
    for a, b in zip(nv_ptr[:-1], nv_ptr[1:]):
        f[..., a:b] *= Q / float(b - a)
    fnew = np.zeros_like(f)
    dt = 0.01
    lib.one_time_step(f.ctypes.data_as(ctypes.c_void_p), fnew.ctypes.data_as(ctypes.c_void_p),
                      *[ctypes.c_int(v) for v in n], ctypes.c_double(0.0), ctypes.c_double(dt),
                      (ctypes.c_double * 1)(0.0))
    inner = tuple(slice(v, nn - v) for v, nn in zip(vmax, n))
    pulled = []
    for k in range(Q):
        off = list(ir.in_offsets[k]) + [0] * (3 - dim)
        pulled.append(f[tuple(slice(v + o, nn - v + o) for v, nn, o in zip(vmax, n, off)) + (k,)])
    out = evaluate(ir, pulled, {"dt": dt, "t": 0.0})
    scale = max(np.abs(fnew[inner]).max(), 1.0)
    err = max(np.abs(out[k] - fnew[inner + (k,)]).max() for k in range(Q))
    assert err <= 1e-13 * scale, (test, err)
