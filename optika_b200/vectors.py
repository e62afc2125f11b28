"""
Input grids.  Mirrors ``optika.vectors.ObjectVectorArray``
(``optika/vectors/_vectors_object.py:134-150``): wavelength, field and pupil
coordinates of the rays entering a :class:`~optika_b200.systems.SequentialSystem`.
"""

from __future__ import annotations
import dataclasses
from . import named as na

__all__ = ["ObjectVectorArray", "PolarizationVectorArray", "SpectralPositionalVectorArray"]


@dataclasses.dataclass(eq=False)
class ObjectVectorArray:
    wavelength: float | na.ScalarArray = 0
    field: na.Cartesian2dVectorArray = 0
    pupil: na.Cartesian2dVectorArray = 0

    @property
    def shape(self) -> dict[str, int]:
        return na.shape_broadcasted(self.wavelength, self.field, self.pupil)

    def copy_shallow(self) -> "ObjectVectorArray":
        return dataclasses.replace(self)


@dataclasses.dataclass(eq=False)
class PolarizationVectorArray:
    """``optika.vectors.PolarizationVectorArray``: s and p components."""

    s: float | na.ScalarArray = 0
    p: float | na.ScalarArray = 0

    @property
    def average(self):
        return (self.s + self.p) / 2


@dataclasses.dataclass(eq=False)
class SpectralPositionalVectorArray:
    """``na.SpectralPositionalVectorArray``: a wavelength and a 2-D position."""

    wavelength: float | na.ScalarArray = 0
    position: na.Cartesian2dVectorArray = 0

    @property
    def shape(self) -> dict[str, int]:
        return na.shape_broadcasted(self.wavelength, self.position)
