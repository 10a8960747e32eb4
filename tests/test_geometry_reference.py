"""
The reference's element / geometry unit tests that do not compare pictures (reference:
tests/test_elements.py:4-28 bounds of the unit shapes, tests/test_geometry.py:35-69 labels and the
dimension check) on the six elements this package implements (cylinders are out of scope).  CPU only.
"""
import pytest


def _elements(lb):
    return [
        (2, lb.Circle([0, 0], 1)), (2, lb.Ellipse([0, 0], [1, 0], [0, 1])),
        (2, lb.Triangle([-1, -1], [0, 2], [2, 0])), (2, lb.Parallelogram([-1, -1], [0, 2], [2, 0])),
        (3, lb.Ellipsoid([0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1])), (3, lb.Sphere([0, 0, 0], 1)),
    ]


BOXES = [(1, {"x": [-2, 2], "label": 3}), (2, {"x": [-2, 2], "y": [-2, 2], "label": 3}),
         (3, {"x": [-2, 2], "y": [-2, 2], "z": [-2, 2], "label": 3})]


@pytest.mark.parametrize("i", range(6))
def test_bounds_of_unit_shapes(i):
    import pylbm_b200 as lb

    dim, element = _elements(lb)[i]
    bounds = element.get_bounds()
    assert list(bounds[0]) == pytest.approx([-1] * dim)
    assert list(bounds[1]) == pytest.approx([1] * dim)


@pytest.mark.parametrize("dim_box,box", BOXES, ids=["box1d", "box2d", "box3d"])
def test_box_and_element_labels(dim_box, box):
    import pylbm_b200 as lb

    assert lb.Geometry({"box": box}).list_of_labels() == [3]
    for dim_element, element in _elements(lb):
        dico = {"box": box, "elements": [element]}
        if dim_element != dim_box:
            with pytest.raises(ValueError):
                lb.Geometry(dico)
        else:
            geom = lb.Geometry(dico)
            assert list(geom.list_of_labels()) == pytest.approx([0, 3])
            assert list(geom.list_of_elements_labels()) == pytest.approx([0])
