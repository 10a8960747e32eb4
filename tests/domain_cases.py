"""Geometries of the reference's domain test module (reference: tests/test_domain.py:13-86), built
with the classes of `mod` (the reference for the fixtures, this package for the test)."""


def domain_cases(mod):
    return [
        {"box": {"x": [0, 1], "label": 0}, "space_step": 0.1, "schemes": [{"velocities": list(range(3))}]},
        {"box": {"x": [0, 2], "y": [0, 1], "label": 0},
         "elements": [mod.Ellipse((0.5, 0.5), (0.25, 0.25), (0.1, -0.1), label=1)],
         "space_step": 0.05, "schemes": [{"velocities": list(range(13))}]},
        {"box": {"x": [0, 2], "y": [0, 1], "label": 0}, "elements": [mod.Circle((0.5, 0.5), 0.2, label=1)],
         "space_step": 0.05, "schemes": [{"velocities": list(range(13))}]},
        {"box": {"x": [0, 1], "y": [0, 1], "label": [0, 1, 2, 3]}, "space_step": 0.1,
         "schemes": [{"velocities": list(range(9))}]},
        {"box": {"x": [0, 3], "y": [0, 1], "label": [0, 1, 0, 2]},
         "elements": [mod.Parallelogram((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), label=0)],
         "space_step": 0.125, "schemes": [{"velocities": list(range(9))}]},
        {"box": {"x": [0, 1], "y": [0, 1], "label": 0},
         "elements": [
             mod.Parallelogram((0.4, 0.3), (0, 0.4), (0.2, 0), label=1),
             mod.Circle((0.4, 0.5), 0.2, label=3),
             mod.Circle((0.6, 0.5), 0.2, label=3),
             mod.Parallelogram((0.45, 0.3), (0, 0.4), (0.1, 0), label=2, isfluid=True),
         ],
         "space_step": 0.025, "schemes": [{"velocities": list(range(9))}]},
        {"box": {"x": [0, 3], "y": [0, 3], "z": [0, 3], "label": list(range(1, 7))},
         "elements": [mod.Ellipsoid((1.5, 1.5, 1.5), [0.5, 0, 0], [0, 0.5, 0], [0, 0, 1], label=0)],
         "space_step": 0.5, "schemes": [{"velocities": list(range(19))}]},
        {"box": {"x": [0, 2], "y": [0, 2], "z": [0, 2], "label": list(range(1, 7))},
         "elements": [mod.Sphere((1, 1, 1), 0.5, label=0)],
         "space_step": 0.5, "schemes": [{"velocities": list(range(19))}]},
        {"box": {"x": [0, 2], "y": [0, 2], "z": [0, 2], "label": list(range(6))}, "space_step": 0.5,
         "schemes": [{"velocities": list(range(19))}]},
    ]
