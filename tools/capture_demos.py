#!/usr/bin/env python
"""
Capture the INPUTS and OUTPUTS of the reference's own demo regression tests as fixtures.

  python tools/capture_demos.py [name ...]     # writes tests/golden/demos/*.pkl, *.npz, MANIFEST.json

The reference pins its `one_time_step` path with 26 demo runs (reference:
tests/test_demo_1d.py:14-72, tests/test_demo_2d.py:10-67, tests/test_demo_3d.py:9-33): every demo
module is imported, `run(dx=1/64, Tf=0.5, generator=g, with_plot=False)` is called and the
conserved moments of the returned Simulation are compared with tests/reference/<test>.h5.

This script runs the UNMODIFIED reference (read-only /root/reference, imported through
tools/refshim/) exactly like those tests, with the Cython generator, and records per test

* `<test>.pkl`  the dictionary the demo handed to `pylbm.Simulation` (cloudpickle; init / boundary
                callables by value, sympy expressions as they are).  The two kinds of reference
                OBJECTS a dictionary holds are replaced by neutral descriptions so that the same
                dictionary can be handed to any implementation: geometric elements become
                `("element", class name, args, kwargs)` and boundary-method classes become
                `("bc", class name)`; tests/demo_fixtures.py puts the classes of the package
                under test back.
* `<test>.npz`  `ref_<moment>`: conserved moments of the reference run at the final time (solid
                cells zeroed as the reference's h5diff plugin does, tests/conftest.py:239-242),
                `h5_<moment>`: the reference's own golden field from tests/reference/<test>.h5
                when that file exists (two 3-D files are missing from the reference checkout),
                `nsteps`, `t`.  Large 3-D fields are stored on every 4th x-plane only.

It cannot run on the GPU box (no /root/reference there); the committed fixtures can.
"""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("PYLBM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "tools", "refshim"), REFERENCE, ROOT, os.path.join(ROOT, "tools")]

OUT = os.path.join(ROOT, "tests", "golden", "demos")
PLANE_STRIDE = 4          # 3-D fields above MAX_FULL values are stored on every 4th x-plane
MAX_FULL = 1 << 17

# (test name = golden file stem, demo directory, module, dx)
TESTS = [
    ("test1D_advection", "1D", "advection", 1.0 / 64),
    ("test1D_advection_reaction", "1D", "advection_reaction", 1.0 / 64),
    ("test1D_burgers", "1D", "burgers", 1.0 / 64),
    ("test1D_euler", "1D", "euler", 1.0 / 64),
    ("test1D_riemann_advection", "1D/riemann_problems", "advection", 1.0 / 64),
    ("test1D_riemann_burgers", "1D/riemann_problems", "burgers", 1.0 / 64),
    ("test1D_riemann_euler", "1D/riemann_problems", "euler", 1.0 / 64),
    ("test1D_riemann_euler_isothermal", "1D/riemann_problems", "euler_isothermal", 1.0 / 64),
    ("test1D_riemann_p_system", "1D/riemann_problems", "p_system", 1.0 / 64),
    ("test1D_riemann_shallow_water", "1D/riemann_problems", "shallow_water", 1.0 / 64),
    ("test2D_advection", "2D", "advection", 1.0 / 64),
    ("test2D_advection_init_f", "2D", "advection_init_f", 1.0 / 64),
    ("test2D_air_conditioning", "2D", "air_conditioning", 1.0 / 64),
    ("test2D_coude", "2D", "bend", 1.0 / 64),
    ("test2D_karman_vortex_street", "2D", "Karman_vortex_street", 1.0 / 64),
    ("test2D_kelvin_Helmoltz", "2D", "Kelvin_Helmoltz", 1.0 / 64),
    ("test2D_lid_driven_cavity", "2D", "lid_driven_cavity", 1.0 / 64),
    ("test2D_orszag_Tang_vortex", "2D", "Orszag_Tang_vortex", 2.0 * np.pi / 64),
    ("test2D_poiseuille", "2D", "Poiseuille", 1.0 / 64),
    ("test2D_poiseuille_vec", "2D", "Poiseuille_vec", 1.0 / 64),
    ("test2D_rayleigh_benard", "2D", "Rayleigh-Benard", 1.0 / 64),
    ("test2D_shallow_water", "2D", "shallow_water", 1.0 / 64),
    ("test3D_advection", "3D", "advection", 1.0 / 64),
    ("test3D_karman", "3D", "Karman", 1.0 / 64),
    ("test3D_lid_cavity", "3D", "lid_cavity", 1.0 / 64),
    ("test3D_poseuille", "3D", "poiseuille", 1.0 / 64),
    # demos that exist in the reference but are not in its test modules (no golden file: compared with
    # the reference run only)
    ("extra2D_lid_driven_cavity_2", "2D", "lid_driven_cavity_2", 1.0 / 64),
    ("extra2D_shallow_water_2", "2D", "shallow_water_2", 1.0 / 64),
    ("extra2D_step", "2D", "step", 1.0 / 64),
]
FINAL_TIME = 0.5


LAST_SIMULATION = []


def _copy_containers(obj):
    """copy of the dict / list / tuple skeleton (leaves are shared): the state of the dictionary at
    the moment it is handed over, whatever the constructor does with it afterwards."""
    if isinstance(obj, dict):
        return {k: _copy_containers(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_copy_containers(v) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_copy_containers(v) for v in obj)
    return obj


def _install_recorders(pylbm):
    """element constructors remember their arguments; Simulation remembers its dictionary."""
    import pylbm.elements as elements

    for name in ("Circle", "Ellipse", "Parallelogram", "Triangle", "Sphere", "Ellipsoid",
                 "CylinderCircle", "CylinderEllipse", "CylinderTriangle", "Parallelepiped"):
        cls = getattr(pylbm, name, None)
        if cls is None:
            continue

        def make(cls=cls, name=name):
            class Recorded(cls):          # same behaviour, plus the constructor arguments
                def __init__(self, *args, **kwargs):
                    self._ctor = (name, _copy_containers(args), _copy_containers(kwargs))
                    super().__init__(*args, **kwargs)
            Recorded.__name__ = name
            Recorded.__qualname__ = name
            return Recorded

        rec = make()
        setattr(pylbm, name, rec)
        if hasattr(elements, name):
            setattr(elements, name, rec)

    original = pylbm.Simulation.__init__

    def recording_init(self, dico, *args, **kwargs):
        self._captured = (_copy_containers(dico), _copy_containers(args), _copy_containers(kwargs))
        original(self, dico, *args, **kwargs)
        LAST_SIMULATION[:] = [self]

    pylbm.Simulation.__init__ = recording_init


def _neutral(dico, pylbm):
    """reference objects -> neutral descriptions (see the module docstring)."""
    from pylbm.boundary import BoundaryMethod

    def conv(obj):
        if isinstance(obj, dict):
            return {k: conv(v) for k, v in obj.items()}
        if isinstance(obj, list):
            return [conv(v) for v in obj]
        if isinstance(obj, tuple):
            return tuple(conv(v) for v in obj)
        if isinstance(obj, type) and issubclass(obj, BoundaryMethod):
            return ("bc", obj.__name__)
        if hasattr(obj, "_ctor"):
            name, args, kwargs = obj._ctor
            return ("element", name, args, kwargs)
        return obj

    return conv(dico)


def _fields(sol):
    """conserved moments with solid cells zeroed (reference: tests/conftest.py:225-262)."""
    domain = sol.domain
    slices = tuple(slice(v, -v) for v in domain.stencil.vmax[: domain.dim])
    solid = domain.in_or_out[slices] != domain.valin
    out = {}
    for key in sol.scheme.consm:
        data = np.array(sol.m[key], dtype=float, copy=True)
        data[solid] = 0.0
        out[str(key)] = data
    return out


def _thin(arr):
    if arr.ndim == 3 and arr.size > MAX_FULL:
        return np.ascontiguousarray(arr[::PLANE_STRIDE])
    return arr


def capture(test, directory, module, dx):
    import cloudpickle
    import pylbm
    from make_golden import convert_h5

    path = os.path.join(REFERENCE, "demo", directory)
    sys.path.append(path)
    try:
        spec = importlib.util.spec_from_file_location(module, os.path.join(path, module + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        del LAST_SIMULATION[:]
        try:
            sol = mod.run(dx, FINAL_TIME, generator="cython", with_plot=False)
        except AttributeError:
            # a demo outside the reference's test modules opens its viewer unconditionally (stub
            # matplotlib here): take the Simulation it built and run the demos' loop ourselves
            if not LAST_SIMULATION:
                raise
            sol = LAST_SIMULATION[0]
            while sol.t < FINAL_TIME:
                sol.one_time_step()
    finally:
        sys.path.remove(path)
    dico, args, kwargs = sol._captured
    record = {
        "test": test, "demo": "demo/%s/%s.py" % (directory, module), "space_step": dx,
        "final_time": FINAL_TIME, "dico": _neutral(dico, pylbm), "sim_args": args, "sim_kwargs": kwargs,
    }
    os.makedirs(OUT, exist_ok=True)
    # helper modules of the demos (e.g. demo/1D/riemann_problems/exact_solvers) are importable while
    # the demo runs: their functions must travel by value as well
    demo_root = os.path.join(REFERENCE, "demo")
    for helper in list(sys.modules.values()):
        if (getattr(helper, "__file__", None) or "").startswith(demo_root):
            try:
                cloudpickle.register_pickle_by_value(helper)
            except ValueError:
                pass
    with open(os.path.join(OUT, test + ".pkl"), "wb") as fh:
        fh.write(cloudpickle.dumps(record, protocol=4))

    arrays = {"nsteps": np.array(sol.nt), "t": np.array(sol.t), "plane_stride": np.array(1)}
    fields = _fields(sol)
    thinned = any(_thin(v) is not v for v in fields.values())
    if thinned:
        arrays["plane_stride"] = np.array(PLANE_STRIDE)
    for key, val in fields.items():
        arrays["ref_" + key] = _thin(val)
    h5 = os.path.join(REFERENCE, "tests", "reference", test + ".h5")
    h5_err = None
    if os.path.exists(h5):
        gold = convert_h5(test + ".h5")
        h5_err = 0.0
        for key, val in fields.items():
            g = gold[key].reshape(val.shape)
            h5_err = max(h5_err, float(np.abs(g - val).max()))
            arrays["h5_" + key] = _thin(g)
    np.savez_compressed(os.path.join(OUT, test + ".npz"), **arrays)
    return {
        "demo": record["demo"], "space_step": dx, "final_time": FINAL_TIME, "nsteps": int(sol.nt),
        "moments": sorted(fields), "shape": list(next(iter(fields.values())).shape),
        "h5_golden": os.path.exists(h5), "reference_run_vs_h5_max_abs": h5_err,
        "plane_stride": int(arrays["plane_stride"]),
    }


def main():
    import pylbm

    _install_recorders(pylbm)
    # the Riemann demos draw the exact solution's wave diagram unconditionally: give the stub
    # pyplot a do-nothing figure
    import matplotlib.pyplot as plt

    class _Anything:
        def __getattr__(self, name):
            return lambda *a, **k: _Anything()

        def __iter__(self):
            return iter(())

    plt.figure = lambda *a, **k: _Anything()
    plt.show = lambda *a, **k: None

    only = set(sys.argv[1:])
    mpath = os.path.join(OUT, "MANIFEST.json")
    manifest = {}
    if os.path.exists(mpath):
        with open(mpath) as fh:
            manifest = json.load(fh)
    manifest.setdefault("reference", "pylbm 0.11.0 (unmodified, /root/reference), Cython generator")
    manifest.setdefault("tests", {})
    for test, directory, module, dx in TESTS:
        if only and test not in only:
            continue
        print("reference run:", test, flush=True)
        manifest["tests"][test] = capture(test, directory, module, dx)
        print("   ", manifest["tests"][test], flush=True)
        with open(mpath, "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
