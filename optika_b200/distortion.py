"""
Distortion models: how a system maps scene coordinates (wavelength, field) onto its sensor.

Mirrors ``optika.distortion`` (``optika/distortion/_distortion.py``) for the consumer of ray
output that SURVEY.md section 8f-4 names: :class:`PolynomialDistortionModel` (``:276-411``), built by
``SequentialSystem.distortion`` from per-field-point means over the pupil that the trace kernel
itself accumulates (``optk_image_t.group_size``).  Plotting is out of scope.
"""

from __future__ import annotations
import dataclasses
import functools
import numpy as np
from . import named as na
from ._polynomial import PolynomialFit
from .vectors import SpectralPositionalVectorArray

__all__ = ["AbstractDistortionModel", "PolynomialDistortionModel"]


@dataclasses.dataclass(eq=False)
class AbstractDistortionModel:
    """``optika.distortion.AbstractDistortionModel`` (``_distortion.py:20-110``)."""

    def distort(self, coordinates: SpectralPositionalVectorArray) -> SpectralPositionalVectorArray:
        raise NotImplementedError

    def undistort(self, coordinates: SpectralPositionalVectorArray) -> SpectralPositionalVectorArray:
        raise NotImplementedError


def _mean(a, axes) -> na.ScalarArray:
    a = na.as_named_array(a)
    present = tuple(ax for ax in axes if ax in a.axes)
    return a.mean(present) if present else a


@dataclasses.dataclass(eq=False)
class PolynomialDistortionModel(AbstractDistortionModel):
    """
    Forward and inverse polynomial fits between scene and sensor coordinates
    (``optika/distortion/_distortion.py:276-425``): :meth:`distort` maps (wavelength, field) to the
    sensor position, :meth:`undistort` is a SEPARATE fit from (wavelength, sensor position) back to
    the field, both of total degree `degree` about the mean of their inputs, over the points
    selected by `where`.
    """

    coordinates_scene: SpectralPositionalVectorArray = None
    coordinates_sensor: na.Cartesian2dVectorArray = None
    axis_wavelength: str = None
    axis_field: tuple = None
    degree: int = 1
    where: object = True

    @property
    def _axis_scene(self) -> tuple:
        return (self.axis_wavelength, *self.axis_field)

    def _fit(self, position_in, position_out) -> PolynomialFit:
        scene = self.coordinates_scene
        inputs = (scene.wavelength, position_in.x, position_in.y)
        shape_ = na.broadcast_shapes(*[na.shape(a) for a in inputs], na.shape(position_out))
        inputs = tuple(na.broadcast_to(na.as_named_array(a), {ax: n for ax, n in shape_.items()
                                                              if ax in self._axis_scene or ax in na.shape(a)})
                       for a in inputs)
        return PolynomialFit(
            inputs=inputs,
            outputs=(position_out.x, position_out.y),
            degree=self.degree,
            center=tuple(_mean(a, self._axis_scene) for a in inputs),  # :389, :405
            where=self.where,
            axes=tuple(ax for ax in shape_ if ax in self._axis_scene),
        )

    @functools.cached_property
    def fit(self) -> PolynomialFit:
        """Scene position -> sensor position (``:383-393``)."""
        return self._fit(self.coordinates_scene.position, self.coordinates_sensor)

    @functools.cached_property
    def fit_inverse(self) -> PolynomialFit:
        """Sensor position -> scene position (``:395-411``)."""
        return self._fit(self.coordinates_sensor, self.coordinates_scene.position)

    def distort(self, coordinates: SpectralPositionalVectorArray) -> SpectralPositionalVectorArray:
        x, y = self.fit(coordinates.wavelength, coordinates.position.x, coordinates.position.y)
        return SpectralPositionalVectorArray(coordinates.wavelength, na.Cartesian2dVectorArray(x, y))

    def undistort(self, coordinates: SpectralPositionalVectorArray) -> SpectralPositionalVectorArray:
        x, y = self.fit_inverse(coordinates.wavelength, coordinates.position.x, coordinates.position.y)
        return SpectralPositionalVectorArray(coordinates.wavelength, na.Cartesian2dVectorArray(x, y))

    @property
    def residual(self) -> na.ScalarArray:
        """``|coordinates_sensor - fit.predictions|`` (what ``plot_residual`` shows, ``:462-464``); NaN where masked."""
        px, py = self.fit.predictions
        dx = na.as_named_array(self.coordinates_sensor.x) - px
        dy = na.as_named_array(self.coordinates_sensor.y) - py
        r = np.sqrt(dx * dx + dy * dy)
        where = na.broadcast_to(na.as_named_array(self.where), r.shape)
        return na.ScalarArray(np.where(where.ndarray, r.ndarray, np.nan), r.axes)
