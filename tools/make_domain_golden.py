#!/usr/bin/env python
"""
Domain fixtures from the UNMODIFIED reference for the geometries of its own domain test module
(reference: tests/test_domain.py:13-86, which only compares pictures): `distance`, `flag`, `in_or_out`
of `pylbm.Domain(case)` for every case this package's elements cover (the CylinderEllipse case is left
out: cylinders are out of scope).  Writes tests/golden/domain_cases.npz.

  python tools/make_domain_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("PYLBM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "tools", "refshim"), REFERENCE, ROOT, os.path.join(ROOT, "tests")]


def main():
    import pylbm
    from domain_cases import domain_cases

    out = {}
    for i, case in enumerate(domain_cases(pylbm)):
        dom = pylbm.Domain(case)
        out["c%d_distance" % i] = dom.distance
        out["c%d_flag" % i] = dom.flag
        out["c%d_in_or_out" % i] = dom.in_or_out
        print(i, dom.distance.shape, int((dom.flag != 999).sum()), "cut links")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "domain_cases.npz"), **out)


if __name__ == "__main__":
    main()
