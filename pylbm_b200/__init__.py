"""
pylbm_b200 -- B200-native lattice Boltzmann time-step engine behind the pylbm API.

Same public names as the reference package (reference: pylbm/__init__.py:24-36)
for everything on the `one_time_step` path: `Simulation`, `Domain`, `Scheme`,
`Stencil`, `Geometry`, the geometric elements and the boundary methods in `bc`.
The only generator is `'cuda'`; there is no CPU fallback.
"""

__version__ = "0.1.0"

from .stencil import Stencil, Velocity  # noqa: F401
from .geometry import Geometry  # noqa: F401
from .elements import Circle, Ellipse, Parallelogram, Triangle, Sphere, Ellipsoid  # noqa: F401
from .domain import Domain, SlabTopology  # noqa: F401
from .scheme import Scheme  # noqa: F401
from . import boundary as bc  # noqa: F401
from .simulation import Simulation  # noqa: F401
from .hdf5 import H5File  # noqa: F401

__all__ = [
    "Simulation", "Domain", "Scheme", "Stencil", "Velocity", "Geometry", "bc",
    "Circle", "Ellipse", "Parallelogram", "Triangle", "Sphere", "Ellipsoid", "SlabTopology", "H5File",
]
