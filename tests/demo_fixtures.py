"""
Loader of the reference-demo fixtures of tests/golden/demos/ (made by tools/capture_demos.py from
the unmodified reference): the dictionary a reference demo hands to `pylbm.Simulation`, with the
classes of the implementation under test put back, and the reference's results.
"""
import json
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEMOS = os.path.join(HERE, "golden", "demos")


def manifest():
    with open(os.path.join(DEMOS, "MANIFEST.json")) as fh:
        return json.load(fh)["tests"]


def demo_names():
    return sorted(manifest())


def _resolve(obj, mod):
    """neutral descriptions -> classes / objects of `mod` (tools/capture_demos.py:_neutral)."""
    if isinstance(obj, dict):
        return {k: _resolve(v, mod) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_resolve(v, mod) for v in obj]
    if isinstance(obj, tuple):
        if len(obj) == 2 and obj[0] == "bc" and isinstance(obj[1], str):
            return getattr(mod.bc, obj[1])
        if len(obj) == 4 and obj[0] == "element" and isinstance(obj[1], str):
            return getattr(mod, obj[1])(*obj[2], **obj[3])
        return tuple(_resolve(v, mod) for v in obj)
    return obj


def load_demo(test, mod=None, generator="cuda"):
    """(dictionary, constructor kwargs, record) of one reference demo test."""
    if mod is None:
        import pylbm_b200 as mod
    with open(os.path.join(DEMOS, test + ".pkl"), "rb") as fh:
        record = pickle.load(fh)
    if "early" in record:         # notebooks: the dictionary was pickled when the Simulation was built
        record["dico"], record["sim_args"], record["sim_kwargs"] = pickle.loads(record["early"])
    dico = _resolve(record["dico"], mod)
    dico["generator"] = generator
    return dico, dict(record["sim_kwargs"]), record


def load_results(test):
    """{'ref': {moment: field}, 'h5': {moment: field} or None, nsteps, plane_stride}."""
    data = np.load(os.path.join(DEMOS, test + ".npz"))
    out = {"ref": {}, "h5": {}, "nsteps": int(data["nsteps"]), "t": float(data["t"]),
           "plane_stride": int(data["plane_stride"])}
    for key in data.files:
        if key.startswith("ref_"):
            out["ref"][key[4:]] = data[key]
        elif key.startswith("h5_"):
            out["h5"][key[3:]] = data[key]
    if not out["h5"]:
        out["h5"] = None
    return out


def final_fields(sim, plane_stride=1):
    """conserved moments with solid cells zeroed, like the reference's h5diff plugin
    (reference: tests/conftest.py:225-262)."""
    domain = sim.domain
    inner = tuple(slice(v, -v) for v in list(domain.stencil.vmax)[: domain.dim])
    solid = domain.in_or_out[inner] != domain.valin
    out = {}
    for key in sim.scheme.consm:
        field = np.array(sim.m[key], dtype=float, copy=True)
        field[solid] = 0.0
        if plane_stride > 1:
            field = field[::plane_stride]
        out[str(key)] = field
    return out


def run_to_final_time(sim, final_time):
    """the loop of every reference demo: `while sol.t < Tf: sol.one_time_step()`."""
    while sim.t < final_time:
        sim.one_time_step()
    return sim
