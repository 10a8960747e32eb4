"""
The reference's Stencil / Velocity unit tests on this package's stencil module (reference:
tests/test_stencil.py:5-62, tests/test_velocity.py:6-32) with the reference's velocity tables D1Q2 ...
D3Q27 (tests/conftest.py:19-163, stored as tests/golden/stencils.json by tools/make_golden.py).  The velocity NUMBERING is part of the parity contract: it
fixes the population order Q of every array and boundary list.  CPU only.
"""
import json
import os
import re

import numpy as np
import pytest

TABLES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stencils.json")))


def _dico(name):
    dim = int(re.search(r"D(\d)", name).group(1))
    table = TABLES[name]
    velocities = [{"velocities": v["num"]} for v in table] if isinstance(table, list) else [{"velocities": table["num"]}]
    return dim, table, {"dim": dim, "schemes": velocities}


@pytest.mark.parametrize("name", sorted(TABLES))
def test_stencil_tables(name):
    import pylbm_b200 as lb

    dim, table, dico = _dico(name)
    assert lb.Stencil.extract_dim(dico) == dim
    stencil = lb.Stencil(dico)
    assert stencil.is_symmetric()
    for bad in (-3, 5):
        with pytest.raises(ValueError):
            stencil.get_symmetric(axis=bad)
    comps = ["vx", "vy", "vz"][:dim]
    all_vel = stencil.get_all_velocities(0)
    for d, c in enumerate(comps):
        assert np.array_equal(all_vel[:, d], getattr(stencil, c)[0])
    if isinstance(table, list):
        for il, entry in enumerate(table):
            for c in comps:
                assert np.array_equal(getattr(stencil, c)[il], entry[c]), (name, il, c)
    else:
        assert np.array_equal(stencil.num[0], table["num"])
        assert np.array_equal(stencil.unum, table["num"])
        assert stencil.unvtot == len(table["num"])
        for c in comps:
            assert np.array_equal(getattr(stencil, c)[0], table[c])
            assert np.array_equal(getattr(stencil, "u" + c), table[c])


def test_velocity_needs_arguments():
    from pylbm_b200.stencil import Velocity

    with pytest.raises(Exception):
        Velocity()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_velocity_symmetries_are_involutions(dim):
    from pylbm_b200.stencil import Velocity

    rng = np.random.default_rng(dim)
    for i in rng.integers(1000, size=100):
        v = Velocity(dim=dim, num=int(i))
        for a in [None] + list(range(dim)):
            vs = v.get_symmetric(axis=a).get_symmetric(axis=a)
            assert vs.v == v.v and vs.num == v.num
        axes = rng.integers(dim, size=5)
        vs = v
        for a in list(axes) + list(axes[::-1]):
            vs = vs.get_symmetric(axis=int(a))
        assert vs.v == v.v and vs.num == v.num
