"""
Input grids.  Mirrors ``optika.vectors.ObjectVectorArray``
(``optika/vectors/_vectors_object.py:134-150``): wavelength, field and pupil
coordinates of the rays entering a :class:`~optika_b200.systems.SequentialSystem`.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na

__all__ = [
    "ObjectVectorArray",
    "PolarizationVectorArray",
    "SpectralPositionalVectorArray",
    "SpectralDirectionalVectorArray",
]


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

    def cell_area(
        self,
        axis_wavelength: str,
        axis_field: tuple[str, str],
        axis_pupil: tuple[str, str],
        field_is_angular: bool = True,
        pupil_is_angular: bool = False,
        factors: bool = False,
    ):
        """
        5-dimensional area of every grid cell: wavelength x field x pupil, in
        mm x (sr or mm^2) x (mm^2 or sr) (``optika/vectors/_vectors_object.py:42-133``).
        Angular grids contribute the solid angle of ``optika.direction`` of their vertices.
        The reference tells angles from lengths by their astropy unit; here the caller
        says which is which.  With ``factors`` the three factors are returned separately.
        """
        from . import _util

        wavelength = na.as_named_array(self.wavelength)
        if axis_wavelength not in wavelength.shape:
            raise ValueError(f"{axis_wavelength=} must be in {wavelength.shape=}")
        for name, v, axes in (("field", self.field, axis_field), ("pupil", self.pupil, axis_pupil)):
            if not set(axes).issubset(na.shape(v)):
                raise ValueError(f"axes {axes} of the {name} grid must be a subset of {na.shape(v)}")
        area_wavelength = wavelength.volume_cell(axis_wavelength)

        def area(v, axes, angular):
            a = _util.direction(v).solid_angle_cell(axes) if angular else v.volume_cell(axes)
            return np.abs(a)

        area_field = area(self.field, axis_field, field_is_angular).cell_centers(axis_wavelength)
        area_pupil = area(self.pupil, axis_pupil, pupil_is_angular).cell_centers((axis_wavelength,) + tuple(axis_field))
        if factors:
            return area_wavelength, area_field, area_pupil
        return area_wavelength * area_field * area_pupil


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


@dataclasses.dataclass(eq=False)
class SpectralDirectionalVectorArray:
    """``na.SpectralDirectionalVectorArray``: a wavelength and a direction (measured efficiencies)."""

    wavelength: float | na.ScalarArray = 0
    direction: na.Cartesian3dVectorArray | float = 0

    @property
    def shape(self) -> dict[str, int]:
        return na.shape_broadcasted(self.wavelength, self.direction)
