"""
Sag profiles: the shape ``z(x, y)`` of an optical surface.

Host-side descriptions with the field names of ``optika.sags`` (reference:
``optika/sags/_flat.py``, ``_spherical.py``, ``_cylindrical.py``, ``_conic.py``,
``_parabolic.py``, ``_toroidal.py``, ``_abc.py``).  These classes hold parameters
only; the arithmetic (intercept, normal, Beer-Lambert attenuation) runs in the
fused CUDA kernel ``optika_b200/csrc/trace_impl.cuh``.  The unit operations
``__call__``, ``normal``, ``intercept`` and ``propagate_rays`` keep the reference
signatures (``optika/sags/_abc.py:48-122``) and are executed on the device by a
one-surface trace restricted to the relevant stages.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na
from . import units as u
from .transformations import AbstractTransformation

__all__ = [
    "AbstractSag",
    "NoSag",
    "SphericalSag",
    "CylindricalSag",
    "ConicSag",
    "ParabolicSag",
    "ToroidalSag",
]


@dataclasses.dataclass(eq=False)
class AbstractSag:
    """Interface of a sag profile (``optika/sags/_abc.py:17-122``)."""

    @property
    def _parameters(self) -> tuple:
        return ()

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(
            *[na.shape(p) for p in self._parameters],
            na.shape(self.transformation),
        )

    # -- unit operations, executed on the device ---------------------------
    def __call__(self, position: na.Cartesian3dVectorArray):
        """Sag ``z(x, y)`` at `position` (``optika/sags/_abc.py:48-61``)."""
        from . import _engine

        return _engine.sag_value(self, position)

    def normal(self, position: na.Cartesian3dVectorArray):
        """Unit normal at `position` (``optika/sags/_abc.py:63-74``)."""
        from . import _engine

        return _engine.sag_normal(self, position)

    def intercept(self, rays):
        """Rays moved to their intercept with this sag (``optika/sags/_abc.py:76-107``)."""
        from . import _engine

        return _engine.sag_intercept(self, rays, attenuate=False)

    def propagate_rays(self, rays):
        """Intercept plus Beer-Lambert attenuation (``optika/sags/_abc.py:109-122``)."""
        from . import _engine

        return _engine.sag_intercept(self, rays, attenuate=True)


@dataclasses.dataclass(eq=False)
class NoSag(AbstractSag):
    """A flat surface, ``z = 0`` (``optika/sags/_flat.py:13-64``)."""

    transformation: None | AbstractTransformation = None


@dataclasses.dataclass(eq=False)
class SphericalSag(AbstractSag):
    """A sphere of the given `radius` of curvature (``optika/sags/_spherical.py:73-191``)."""

    radius: float | na.ScalarArray = np.inf
    transformation: None | AbstractTransformation = None

    @property
    def curvature(self):
        return 1 / u.length(self.radius)

    @property
    def _parameters(self):
        return (self.radius,)


@dataclasses.dataclass(eq=False)
class CylindricalSag(AbstractSag):
    """A cylinder curved along x (``optika/sags/_cylindrical.py:16-160``)."""

    radius: float | na.ScalarArray = np.inf
    transformation: None | AbstractTransformation = None

    @property
    def _parameters(self):
        return (self.radius,)


@dataclasses.dataclass(eq=False)
class ConicSag(AbstractSag):
    """A conic section of revolution (``optika/sags/_conic.py:174-252``)."""

    radius: float | na.ScalarArray = np.inf
    conic: float | na.ScalarArray = 0
    transformation: None | AbstractTransformation = None

    @property
    def _parameters(self):
        return (self.radius, self.conic)


@dataclasses.dataclass(eq=False)
class ParabolicSag(AbstractSag):
    """A paraboloid of the given `focal_length` (``optika/sags/_parabolic.py:14-158``)."""

    focal_length: float | na.ScalarArray = np.inf
    transformation: None | AbstractTransformation = None

    @property
    def radius(self):
        return 2 * u.length(self.focal_length)

    @property
    def conic(self) -> int:
        return -1

    @property
    def _parameters(self):
        return (self.focal_length,)


@dataclasses.dataclass(eq=False)
class ToroidalSag(AbstractSag):
    """A toroid: minor `radius`, major `radius_of_rotation` (``optika/sags/_toroidal.py:14-88``)."""

    radius: float | na.ScalarArray = np.inf
    radius_of_rotation: float | na.ScalarArray = 0
    transformation: None | AbstractTransformation = None

    @property
    def _parameters(self):
        return (self.radius, self.radius_of_rotation)
