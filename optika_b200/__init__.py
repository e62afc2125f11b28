"""
optika_b200: a B200-native (sm_100a) engine for the sequential-raytrace hot path
of sun-data/optika, behind optika's own Python API.

The modules mirror the reference's names for the parts on the hot path
(``surfaces``, ``sags``, ``rulings``, ``apertures``, ``materials``, ``sensors``,
``rays``, ``vectors``, ``propagators``, ``systems``); ``named`` (``na``),
``transformations`` and ``units`` (``u``) stand in for the third-party
``named_arrays`` / ``astropy.units`` the reference builds on.  All ray
arithmetic runs in hand-written CUDA behind the C ABI of ``include/optk.h``;
there is no CPU fallback.
"""

from . import units
from . import named
from . import transformations
from . import chemicals
from . import rays
from . import vectors
from . import sags
from . import apertures
from . import rulings
from . import materials
from . import surfaces
from . import sensors
from . import propagators
from . import systems
from ._util import direction, angles, shape

__all__ = [
    "units",
    "named",
    "transformations",
    "chemicals",
    "rays",
    "vectors",
    "sags",
    "apertures",
    "rulings",
    "materials",
    "surfaces",
    "sensors",
    "propagators",
    "systems",
    "direction",
    "angles",
    "shape",
]
