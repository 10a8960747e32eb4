"""
The reference's Domain unit tests on the sparse Domain of this package (reference:
tests/domain/test_domain2D.py:10-113 with the golden files tests/domain/data/*.npz, re-saved by
tools/make_golden.py into tests/golden/domain/; tests/domain/test_domain1D.py:19-126 whose expected
arrays are spelled out in the test).  `distance`, `flag` and `in_or_out` feed the boundary lists, so
they are part of the parity contract: compared with `==` (the reference uses allclose(1e-16) for
distance).  CPU only.
"""
import copy
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "domain")
VALIN, VALOUT = 999, -1

DOM2D = {"box": {"x": [0, 1], "y": [0, 2], "label": 0}, "space_step": 0.25,
         "schemes": [{"velocities": list(range(5))}]}
DOM1D = {"box": {"x": [0, 1], "label": 0}, "space_step": 0.25, "schemes": [{"velocities": list(range(3))}]}


def _elements(name, lb):
    whole = lb.Parallelogram([0.0, 0.0], [1.0, 0], [0.0, 2.0], label=20)
    return {
        "simple_domain": [],
        "rectangle": [lb.Parallelogram([0.23, 0.73], [0.5, 0], [0.0, 0.5], label=10)],
        "fluid_rectangle": [whole, lb.Parallelogram([0.23, 0.73], [0.5, 0], [0.0, 0.5], label=10, isfluid=True)],
        "circle": [lb.Circle([0.5, 1.0], 0.5, label=10)],
        "fluid_circle": [whole, lb.Circle([0.5, 1.0], 0.5, label=10, isfluid=True)],
        "triangle": [lb.Triangle([0.23, 0.73], [0.5, 0], [0.0, 0.5], label=10)],
        "fluid_triangle": [whole, lb.Triangle([0.23, 0.73], [0.5, 0], [0.0, 0.5], label=10, isfluid=True)],
    }[name]


@pytest.mark.parametrize("name", ["simple_domain", "rectangle", "fluid_rectangle", "circle", "fluid_circle",
                                  "triangle", "fluid_triangle"])
def test_domain2d_against_reference_golden(name):
    import pylbm_b200 as lb

    dico = copy.deepcopy(DOM2D)
    elements = _elements(name, lb)
    if elements:
        dico["elements"] = elements
    dom = lb.Domain(dico)
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert np.array_equal(dom.in_or_out, ref["in_or_out"])
    assert np.array_equal(dom.flag, ref["flag"])
    assert np.array_equal(dom.distance, ref["distance"])


def test_domain2d_grid_and_labels():
    import pylbm_b200 as lb

    dico = copy.deepcopy(DOM2D)
    dico["box"]["label"] = [0, 1, 2, 0]
    dom = lb.Domain(dico)
    assert dom.dx == 0.25
    assert np.all(dom.x_halo == np.linspace(-0.125, 1.125, 6))
    assert np.all(dom.y_halo == np.linspace(-0.125, 2.125, 10))


def test_domain1d_one_scheme():
    import pylbm_b200 as lb

    dico = copy.deepcopy(DOM1D)
    dico["box"] = {"x": [0, 1], "label": [0, 1]}
    dom = lb.Domain(dico)
    assert dom.shape_halo == [6] and dom.shape_in == [4] and dom.dx == 0.25
    assert np.all(dom.x_halo == np.linspace(-0.125, 1.125, 6))

    dom = lb.Domain(DOM1D)
    in_or_out = VALIN * np.ones(6)
    in_or_out[[0, -1]] = VALOUT
    assert np.all(dom.in_or_out == in_or_out)
    distance = VALIN * np.ones((3, 6))
    distance[(1, 2), (-2, 1)] = 0.5
    assert np.all(dom.distance == distance)
    flag = VALIN * np.ones((3, 6), dtype=int)
    flag[(1, 2), (-2, 1)] = 0
    assert np.all(dom.flag == flag)


@pytest.mark.parametrize("labels", [None, (1, 2)])
def test_domain1d_two_schemes(labels):
    """D1Q3 + D1Q5: two ghost cells per side, links of length 2 cut at 0.25 and 0.75."""
    import pylbm_b200 as lb

    dico = copy.deepcopy(DOM1D)
    dico["schemes"].append({"velocities": list(range(5))})
    lleft = lright = 0
    if labels:
        lleft, lright = labels
        dico["box"] = {"x": [0, 1], "label": [lleft, lright]}
    dom = lb.Domain(dico)
    in_or_out = VALIN * np.ones(8)
    in_or_out[[0, 1, -2, -1]] = VALOUT
    assert np.all(dom.in_or_out == in_or_out)
    distance = VALIN * np.ones((5, 8))
    distance[(1, 2), (-3, 2)] = 0.5
    distance[(3, 4), (-3, 2)] = 0.25
    distance[(3, 4), (-4, 3)] = 0.75
    assert np.all(dom.distance == distance)
    flag = VALIN * np.ones((5, 8), dtype=int)
    flag[(2, 4, 4), (2, 2, 3)] = lleft
    flag[(1, 3, 3), (-3, -3, -4)] = lright
    assert np.all(dom.flag == flag)


@pytest.mark.parametrize("index", range(9))
def test_domain_cases_of_the_reference_test_module(index):
    """geometries of the reference's tests/test_domain.py (which only compares pictures): ellipse and
    circle under D2Q13 (two ghost layers), labelled box faces, overlapping solid and fluid elements,
    ellipsoid / sphere / plain box under D3Q19 -- fixtures from the unmodified reference
    (tools/make_domain_golden.py)."""
    import pylbm_b200 as lb
    from domain_cases import domain_cases

    ref = np.load(os.path.join(os.path.dirname(GOLDEN), "domain_cases.npz"))
    dom = lb.Domain(domain_cases(lb)[index])
    assert np.array_equal(dom.in_or_out, ref["c%d_in_or_out" % index])
    assert np.array_equal(dom.flag, ref["c%d_flag" % index])
    assert np.array_equal(dom.distance, ref["c%d_distance" % index])
