#!/usr/bin/env python
"""
SASS summary of the fused kernel of a workload (cuobjdump -sass on the in-tree kernel library):
instruction count per instantiation, the mnemonics that matter for an HBM-bound fp64 stencil
(LDG / STG / DFMA / DADD / DMUL / MUFU / F2F / BAR / local-memory LDL/STL = spills), registers.

    python tools/sass_summary.py [case[:key=value,...] ...] > profiles/r02_sass_fused.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = [("lid_cavity_d3q19", dict(n=512), "f64", "f64"), ("lid_cavity_d3q19", dict(n=512), "f64", "f64", True),
           ("lid_cavity_d3q19", dict(n=512), "f32", "f64"),
           ("lid_cavity_d3q19", dict(n=512), "f32", "f32"), ("karman_d2q9", dict(nx=4096, ny=1024), "f64", "f64"),
           ("channel_sphere_d3q27", dict(nx=512, ny=256, nz=256), "f64", "f64")]
WATCH = ["LDG", "STG", "DFMA", "DADD", "DMUL", "FFMA", "FADD", "FMUL", "MUFU", "F2F", "IMAD", "IADD3", "LOP3",
         "ISETP", "BAR", "LDS", "STS", "ATOMS", "LDL", "STL", "BRA"]


def summarize(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\n\s*REG:(\d+)", res))
    out = []
    for block in sass.split("Function : ")[1:]:
        name = block.split("\n", 1)[0].strip()
        if "lbmk_kernel_one_time_step" not in name:
            continue
        counts, total = collections.Counter(), 0
        variants = collections.Counter()
        for line in block.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            total += 1
            counts[op.split(".")[0]] += 1
            if op.startswith(("LDG", "STG", "F2F")):
                variants[op] += 1
        out.append((name, total, counts, variants, regs.get(name)))
    return out


def main():
    from pylbm_b200 import cases
    from pylbm_b200.scheme import Scheme
    from pylbm_b200.simulation import build_kernel_library

    print("# SASS summary of the fused kernel `lbmk_kernel_one_time_step<MODE, TASKS>` (cuobjdump -sass, sm_100a)\n")
    print("Static instruction counts of the whole kernel body (one thread = one cell; the image tail and the\n"
          "task prologue are rarely executed).  No LDL/STL = no register spills.  Not a tensor-core / TMA kernel by\n"
          "design (every population value is read once: nothing to stage or reuse), so UTCMMA / UTMALDG do not\n"
          "apply; the tells here are coalesced `LDG.E.64.CONSTANT` / `STG.E.64.STRONG.GPU` per population and DFMA.\n")
    for entry in DEFAULT:
        name, kw, storage, compute = entry[:4]
        aa = len(entry) > 4
        scheme = Scheme(cases.CASES[name](**kw))
        _, path, _ = build_kernel_library(scheme, storage=storage, compute=compute, aa=aa)
        print("## %s %s, storage %s, arithmetic %s%s (`%s`)\n" % (
            name, kw, storage, compute, ", library for in-place streaming (explicit FMA lowering)" if aa else "",
            os.path.basename(path)))
        print("| instantiation | regs | instr | " + " | ".join(WATCH) + " |")
        print("|---|---|---|" + "---|" * len(WATCH))
        for fn, total, counts, variants, regs in summarize(path):
            tag = re.search(r"ILi(\d)ELb(\d)E", fn)
            modes = {"0": "two arrays", "1": "two arrays + fused walls", "2": "in-place even step",
                     "3": "in-place even step + fused walls"}
            label = "%s%s" % (modes.get(tag.group(1), tag.group(1)), ", task table" if tag.group(2) == "1" else "") if tag else fn[:40]
            print("| %s | %s | %d | " % (label, regs or "?", total) + " | ".join(str(counts.get(w, 0)) for w in WATCH) + " |")
            if tag and tag.group(1) == "0" and tag.group(2) == "0":
                keep = ", ".join("%s x%d" % kv for kv in sorted(variants.items()))
        print("\nmemory / conversion variants of the plain instantiation: %s\n" % keep)


if __name__ == "__main__":
    main()
