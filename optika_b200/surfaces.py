"""
Optical surfaces.

Mirrors ``optika.surfaces.Surface`` (``optika/surfaces.py:282-395``): a sag, a
material, an aperture, optional rulings and a rigid transformation.  The
operator ``propagate_rays`` (``optika/surfaces.py:123-198``) is not evaluated
here: it is lowered to one ``optk_surface_t`` record and executed by the fused
CUDA kernel.
"""

from __future__ import annotations
import dataclasses
from . import named as na
from . import sags as _sags
from . import materials as _materials
from .transformations import AbstractTransformation

__all__ = ["AbstractSurface", "Surface"]


@dataclasses.dataclass(eq=False)
class AbstractSurface:
    """Interface of an optical surface (``optika/surfaces.py:35-198``)."""

    @property
    def is_stop(self) -> bool:
        return self.is_field_stop or self.is_pupil_stop

    @property
    def shape(self) -> dict[str, int]:
        # optika/surfaces.py:386-395
        return na.broadcast_shapes(
            na.shape(self.sag),
            na.shape(self.material),
            na.shape(self.aperture),
            na.shape(self.rulings),
            na.shape(self.transformation),
        )

    def propagate_rays(self, rays):
        """
        Refract, reflect and/or diffract `rays` off this surface
        (``optika/surfaces.py:123-198``), on the device.
        """
        from . import propagators

        return propagators.propagate_rays([self], rays)


@dataclasses.dataclass(eq=False)
class Surface(AbstractSurface):
    """A single optical interface (``optika/surfaces.py:282-395``)."""

    name: None | str = None
    sag: None | _sags.AbstractSag = None
    material: None | _materials.AbstractMaterial = None
    aperture: object = None
    aperture_mechanical: object = None
    rulings: object = None
    is_field_stop: bool = False
    is_pupil_stop: bool = False
    transformation: None | AbstractTransformation = None
    kwargs_plot: None | dict = None

    def __post_init__(self):
        # optika/surfaces.py:380-384
        if self.sag is None:
            self.sag = _sags.NoSag()
        if self.material is None:
            self.material = _materials.Vacuum()
