#!/usr/bin/env python
"""
Capture the Simulations of the reference's tutorial notebooks as fixtures (same format as
tools/capture_demos.py).  The reference's CI executes notebooks/*.ipynb with `pytest --nbval-lax`
(.github/workflows/ci.yml:74): every code cell is run here, in order, in one namespace, with the
UNMODIFIED reference (matplotlib / pylab replaced by do-nothing stubs, IPython magics dropped, a cell
that fails on plotting is skipped).  Every `pylbm.Simulation` a notebook builds and steps is recorded
with the dictionary it received and the conserved moments it holds when the notebook ends
-> tests/golden/demos/nb<NN>_<k>.pkl / .npz, listed in MANIFEST.json next to the demos.

  python tools/capture_notebooks.py [notebook stem ...]
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import capture_demos as cd          # noqa: E402  (sets sys.path for the reference and the shims)

NOTEBOOKS = os.path.join(cd.REFERENCE, "notebooks")
MAX_CELLS = 1 << 22                 # skip simulations whose fields would not be small fixtures
# Simulations without a defined reference result: the source term of the second Simulation of
# 08_advection_reaction depends on X; the reference's generated Cython kernel reads an uninitialised
# local `xx` there and its NumPy kernel raises NameError (algorithm/base.py:590-593 leaves the
# statements that define xx commented out)
SKIP = {"nb08_1"}


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))

    def __getitem__(self, key):
        return _Anything()

    def __setitem__(self, key, value):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _stub_module(name):
    mod = types.ModuleType(name)
    mod.__getattr__ = lambda attr: _Anything()
    sys.modules[name] = mod
    return mod


def main():
    import cloudpickle
    import pylbm

    for name in ("pylab", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1", "IPython", "IPython.display"):
        _stub_module(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import pylbm.viewer.matplotlib_viewer as viewer

    viewer.plt = sys.modules["matplotlib.pyplot"]
    for cls in (pylbm.Geometry, pylbm.Stencil, pylbm.Domain):
        cls.visualize = lambda self, *a, **k: _Anything()

    cd._install_recorders(pylbm)
    built = []
    original = pylbm.Simulation.__init__

    def remember(self, dico, *args, **kwargs):
        original(self, dico, *args, **kwargs)
        built.append(self)
        # pickle NOW: the callables of the dictionary read notebook globals (rhoo, vup, ...) that later
        # cells redefine
        for helper in list(sys.modules.values()):
            if (getattr(helper, "__file__", None) or "").startswith(NOTEBOOKS):
                cloudpickle.register_pickle_by_value(helper)
        captured, cargs, ckwargs = self._captured
        try:
            self._early = cloudpickle.dumps((cd._neutral(captured, pylbm), cargs, ckwargs), protocol=4)
        except Exception as exc:
            self._early = None
            print("   dictionary cannot be pickled (%s)" % exc)

    pylbm.Simulation.__init__ = remember

    mpath = os.path.join(cd.OUT, "MANIFEST.json")
    manifest = json.load(open(mpath))
    only = set(sys.argv[1:])
    for fname in sorted(os.listdir(NOTEBOOKS)):
        if not fname.endswith(".ipynb"):
            continue
        stem = fname[:-6]
        if only and stem not in only:
            continue
        cells = [c for c in json.load(open(os.path.join(NOTEBOOKS, fname)))["cells"] if c["cell_type"] == "code"]
        del built[:]
        namespace = {"__name__": "__notebook__"}
        print("notebook:", stem, len(cells), "code cells", flush=True)
        for i, cell in enumerate(cells):
            src = "".join(cell["source"])
            src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("%", "!")))
            try:
                exec(compile(src, "%s[cell %d]" % (stem, i), "exec"), namespace)
            except Exception as exc:          # plotting on stubs, mostly
                print("   cell %d skipped: %s: %s" % (i, type(exc).__name__, str(exc)[:100]), flush=True)
        for k, sol in enumerate(built):
            test = "nb%s_%d" % (stem[:2], k)
            cells_count = int(np.prod(sol.domain.shape_in))
            if test in SKIP or sol.nt == 0 or cells_count * len(sol.scheme.consm) > MAX_CELLS:
                print("   %s: not recorded (nt=%d, %d cells)" % (test, sol.nt, cells_count))
                continue
            dico, args, kwargs = sol._captured
            if sol._early is None:
                continue
            record = {
                "test": test, "demo": "notebooks/%s (Simulation #%d)" % (fname, k), "space_step": float(sol.domain.dx),
                # `while sol.t < final_time: sol.one_time_step()` then performs exactly sol.nt steps
                "final_time": float(sol.t - 0.5 * sol.dt),
                # (dictionary, args, kwargs) pickled when the Simulation was built
                "early": sol._early,
            }
            blob = cloudpickle.dumps(record, protocol=4)
            with open(os.path.join(cd.OUT, test + ".pkl"), "wb") as fh:
                fh.write(blob)
            fields = cd._fields(sol)
            thinned = any(cd._thin(v) is not v for v in fields.values())
            arrays = {"nsteps": np.array(sol.nt), "t": np.array(sol.t),
                      "plane_stride": np.array(cd.PLANE_STRIDE if thinned else 1)}
            for key, val in fields.items():
                arrays["ref_" + key] = cd._thin(val)
            np.savez_compressed(os.path.join(cd.OUT, test + ".npz"), **arrays)
            manifest["tests"][test] = {
                "demo": record["demo"], "space_step": record["space_step"], "final_time": record["final_time"],
                "nsteps": int(sol.nt), "moments": sorted(fields), "shape": list(next(iter(fields.values())).shape),
                "h5_golden": False, "reference_run_vs_h5_max_abs": None, "plane_stride": int(arrays["plane_stride"]),
                "generator": str(dico.get("generator", "numpy")),
            }
            print("   ", test, manifest["tests"][test], flush=True)
        with open(mpath, "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
