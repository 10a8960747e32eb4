#!/usr/bin/env python
"""
Generate the golden fixtures of tests/golden/ by running the UNMODIFIED reference
(read-only at /root/reference, imported through tools/refshim/) in the build container.

  python tools/make_golden.py            # writes tests/golden/*.npz and tests/golden/MANIFEST.json

Two kinds of fixtures:

1. `ref_<case>.npz` -- for every parity workload of tests/conftest.py (the dictionaries of
   pylbm_b200/cases.py built with `mod=pylbm`, seeded perturbed initial state): the reference's
   boundary lists (istore, iload*, s, rhs, ilabel, distance per method), M / invM, and the
   conserved moments after NSTEPS steps of the reference's Cython generator.
2. `h5_<demo>.npz` -- the reference's OWN golden fields (tests/reference/*.h5, dx = 1/64,
   Tf = 0.5, solid cells zeroed; reference: tests/conftest.py:225-262, test_demo_2d.py:10-67),
   converted with a small HDF5 reader, for the demos whose dictionaries pylbm_b200/cases.py
   reproduces (lid_driven_cavity, Karman_vortex_street, shallow_water).

This script cannot run on the GPU box (no /root/reference there); the committed .npz can.
"""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("PYLBM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "tools", "refshim"), REFERENCE, ROOT, os.path.join(ROOT, "tests")]

NSTEPS = 50
OUT = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------------------
# minimal HDF5 reader (superblock v0, v1 object headers, contiguous f8 datasets)
# ---------------------------------------------------------------------------
class H5Lite:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert self.b[8] == 0, "only superblock version 0 is supported"
        self.O, self.L = self.b[13], self.b[14]
        pos = 24 + 4 * self.O          # after base, free-space, eof, driver addresses
        # root symbol table entry
        _, self.root_header, cache = self._entry(pos)
        assert cache is not None, "root group without cached symbol table"
        self.btree, self.heap = cache

    def _u(self, pos, size):
        return int.from_bytes(self.b[pos : pos + size], "little")

    def _entry(self, pos):
        O = self.O
        name_off = self._u(pos, O)
        header = self._u(pos + O, O)
        ctype = self._u(pos + 2 * O, 4)
        scratch = pos + 2 * O + 8
        cache = (self._u(scratch, O), self._u(scratch + O, O)) if ctype == 1 else None
        return name_off, header, cache

    def _heap_name(self, heap_addr, offset):
        assert self.b[heap_addr : heap_addr + 4] == b"HEAP"
        data = self._u(heap_addr + 8 + 2 * self.L, self.O)
        end = self.b.index(b"\0", data + offset)
        return self.b[data + offset : end].decode()

    def _walk(self, node, heap, out):
        sig = self.b[node : node + 4]
        if sig == b"TREE":
            level = self.b[node + 5]
            used = self._u(node + 6, 2)
            pos = node + 8 + 2 * self.O
            for i in range(used):
                child = self._u(pos + self.L + i * (self.L + self.O), self.O)
                self._walk(child, heap, out)
        elif sig == b"SNOD":
            nsym = self._u(node + 6, 2)
            size = 2 * self.O + 24
            for i in range(nsym):
                name_off, header, _ = self._entry(node + 8 + i * size)
                out[self._heap_name(heap, name_off)] = header
        else:
            raise ValueError("unexpected node %r" % sig)

    def _messages(self, header):
        assert self.b[header] == 1, "only version 1 object headers"
        nmsg = self._u(header + 2, 2)
        size = self._u(header + 8, 4)
        blocks = [(header + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize = self._u(pos, 2), self._u(pos + 2, 2)
                body = pos + 8
                if mtype == 0x10:
                    blocks.append((self._u(body, self.O), self._u(body + self.O, self.L)))
                msgs.append((mtype, body, msize))
                pos = body + msize
        return msgs

    def datasets(self):
        names = {}
        self._walk(self.btree, self.heap, names)
        out = {}
        for name, header in names.items():
            shape = addr = None
            for mtype, body, msize in self._messages(header):
                if mtype == 0x1:
                    version, rank = self.b[body], self.b[body + 1]
                    start = body + (8 if version == 1 else 4)
                    shape = tuple(self._u(start + i * self.L, self.L) for i in range(rank))
                elif mtype == 0x3:
                    assert (self.b[body] & 0x0F) == 1 and self._u(body + 4, 4) == 8, "expected f8 data"
                elif mtype == 0x8:
                    assert self.b[body] == 3 and self.b[body + 1] == 1, "expected contiguous layout v3"
                    addr = self._u(body + 2, self.O)
            if shape is not None and addr is not None:
                count = int(np.prod(shape)) if shape else 1
                out[name] = np.frombuffer(self.b, dtype="<f8", count=count, offset=addr).reshape(shape).copy()
        return out


def convert_h5(name):
    """reference golden file -> {moment: array[x, y(, z)]} (datasets are stored transposed:
    reference hdf5.py:168, 204-206)."""
    data = H5Lite(os.path.join(REFERENCE, "tests", "reference", name)).datasets()
    fields = {}
    for key, arr in data.items():
        if key.startswith("x_"):
            fields[key] = arr
        else:
            fields[key] = np.ascontiguousarray(arr.T)
    return fields


# ---------------------------------------------------------------------------
# reference runs of the parity workloads
# ---------------------------------------------------------------------------
def reference_fixture(case, kw, nsteps=NSTEPS):
    import pylbm
    from pylbm_b200 import cases

    sim = pylbm.Simulation(cases.CASES[case](mod=pylbm, perturb=0, generator="cython", **kw))
    out = {}
    for i, method in enumerate(sim.bc.methods):
        pre = "bc%d_" % i
        out[pre + "name"] = np.array(type(method).__name__)
        out[pre + "istore"] = method.istore
        for j, il in enumerate(method.iload):
            out[pre + "iload%d" % j] = il
        out[pre + "rhs"] = method.rhs.copy()
        out[pre + "ilabel"] = method.ilabel
        out[pre + "distance"] = method.distance
        if hasattr(method, "s"):
            out[pre + "s"] = method.s
    out["nmethods"] = np.array(len(sim.bc.methods))
    params = list(sim.scheme.param.items())
    out["M"] = np.array(sim.scheme.M.subs(params).tolist(), dtype=float)
    out["invM"] = np.array(sim.scheme.invM.subs(params).tolist(), dtype=float)
    out["in_or_out"] = sim.domain.in_or_out
    for key in sim.scheme.consm:
        out["m0_" + str(key)] = sim.m[key].copy()
    for _ in range(nsteps):
        sim.one_time_step()
    for key in sim.scheme.consm:
        out["m_" + str(key)] = sim.m[key].copy()
    out["nsteps"] = np.array(nsteps)
    return out


def main():
    from conftest import PARITY_CASES, case_id

    os.makedirs(OUT, exist_ok=True)
    manifest = {"reference": "pylbm 0.11.0 (unmodified, /root/reference) with the Cython generator",
                "nsteps": NSTEPS, "files": {}}
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for case, kw in PARITY_CASES:
        name = "ref_%s.npz" % case_id(case, kw)
        manifest["files"][name] = {"case": case, "kwargs": kw, "perturb": 0}
        if only is not None and only not in name and os.path.exists(os.path.join(OUT, name)):
            continue
        print("reference run:", name, flush=True)
        np.savez_compressed(os.path.join(OUT, name), **reference_fixture(case, kw))
    for h5, case, kw, steps in [
        ("test2D_lid_driven_cavity.h5", "lid_cavity_d2q9", dict(n=64), 32),
        ("test2D_karman_vortex_street.h5", "karman_d2q9", dict(nx=128, ny=64, radius=1.0 / 32, cx=0.3), 32),
        ("test2D_shallow_water.h5", "shallow_water_d2q4", dict(n=128), None),
        ("test2D_rayleigh_benard.h5", "rayleigh_benard", dict(nx=128, ny=64), None),
    ]:
        fields = convert_h5(h5)
        name = "h5_" + h5.replace(".h5", ".npz")
        np.savez_compressed(os.path.join(OUT, name), **fields)
        manifest["files"][name] = {"case": case, "kwargs": kw, "source": "tests/reference/" + h5,
                                   "final_time": 0.5, "space_step": 1.0 / 64}
        print("converted:", name, {k: v.shape for k, v in fields.items()})
    # the reference's domain goldens (tests/domain/data/*.npz: distance, flag, in_or_out of seven small
    # 2-D geometries, tests/domain/test_domain2D.py:10-19) re-saved compressed
    src = os.path.join(REFERENCE, "tests", "domain", "data")
    dst = os.path.join(OUT, "domain")
    os.makedirs(dst, exist_ok=True)
    for fname in sorted(os.listdir(src)):
        if fname.endswith(".npz"):
            data = np.load(os.path.join(src, fname))
            np.savez_compressed(os.path.join(dst, fname), **{k: data[k] for k in data.files})
            manifest["files"]["domain/" + fname] = {"source": "tests/domain/data/" + fname}
            print("copied:", fname, {k: data[k].shape for k in data.files})
    # the reference's velocity tables D1Q2 ... D3Q27 (tests/conftest.py:19-163, the `_schemes` dict of its
    # stencil fixtures), read from the source with ast (importing that conftest needs h5py)
    import ast

    tree = ast.parse(open(os.path.join(REFERENCE, "tests", "conftest.py")).read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "_schemes":
            table = eval(compile(ast.Expression(node.value), "conftest", "eval"), {"list": list, "range": range})
            with open(os.path.join(OUT, "stencils.json"), "w") as fh:
                json.dump(table, fh, indent=0, sort_keys=True)
            manifest["files"]["stencils.json"] = {"source": "tests/conftest.py:19-163"}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
