"""
Box + elements description (host-side setup).

Mirror of pylbm.geometry.Geometry / get_box (reference: pylbm/geometry.py:22-143,
322-338): `dim`, `bounds`, `box_label` (2*dim labels: x-, x+, y-, y+, z-, z+),
`list_elem`, `list_of_labels()`.
"""

import numpy as np

__all__ = ["Geometry", "get_box"]


def get_box(dico):
    """dimension and bounds of the box of a dictionary."""
    try:
        box = dico["box"]
    except KeyError:
        raise KeyError("'box' key not found in the geometry definition")
    if "x" not in box:
        raise KeyError("'x' interval not found in the box definition of the geometry")
    bounds = [box["x"]]
    if box.get("y", None) is not None:
        bounds.append(box["y"])
        if box.get("z", None) is not None:
            bounds.append(box["z"])
    return len(bounds), np.asarray(bounds, dtype="f8")


class Geometry:
    def __init__(self, dico, need_validation=True):
        self.dim, self.bounds = get_box(dico)
        lab = dico["box"].get("label", -1)
        if isinstance(lab, (int, np.integer)):
            self.box_label = [int(lab)] * 2 * self.dim
        elif isinstance(lab, (list, tuple)):
            if len(lab) != 2 * self.dim:
                raise ValueError("The list label of the box has the wrong size (must be 2*dim)")
            self.box_label = list(lab)
        else:
            raise ValueError("The labels of the box must be an integer or a list")
        self.list_elem = []
        for elem in dico.get("elements", None) or []:
            self.add_elem(elem)

    def add_elem(self, elem):
        if elem.dim != self.dim:
            raise ValueError("Element must have the same dimension of the box")
        self.list_elem.append(elem)

    def list_of_elements_labels(self):
        labels = np.empty(0)
        for elem in self.list_elem:
            labels = np.union1d(labels, elem.label)
        return labels

    def list_of_labels(self):
        return np.union1d(np.unique(self.box_label), self.list_of_elements_labels())

    def __repr__(self):
        return "Geometry(dim={}, bounds={}, labels={}, {} element(s))".format(
            self.dim, self.bounds.tolist(), self.box_label, len(self.list_elem)
        )
