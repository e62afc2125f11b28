"""
Radiometric models fitted to traced rays: vignetting and effective area.

Mirrors ``optika.radiometry`` for the consumers of ray output that SURVEY.md section 8f-4 names:
:class:`PolynomialVignettingModel` (``optika/radiometry/_vignetting.py:99-180``) and
:class:`InterpolatedEffectiveAreaModel` (``optika/radiometry/_effective_area.py:37-102``), built by
``SequentialSystem.vignetting`` / ``area_effective`` from per-field-point sums over the pupil that the
trace kernel accumulates itself.  Plotting is out of scope.
"""

from __future__ import annotations
import dataclasses
import functools
import numpy as np
from . import named as na
from ._polynomial import PolynomialFit
from .distortion import _mean
from .vectors import SpectralPositionalVectorArray

__all__ = [
    "AbstractVignettingModel",
    "PolynomialVignettingModel",
    "AbstractEffectiveAreaModel",
    "InterpolatedEffectiveAreaModel",
]


@dataclasses.dataclass(eq=False)
class AbstractVignettingModel:
    """``optika.radiometry.AbstractVignettingModel`` (``_vignetting.py:18-60``)."""

    def __call__(self, coordinates: SpectralPositionalVectorArray) -> na.ScalarArray:
        raise NotImplementedError

    def inverse(self, coordinates: SpectralPositionalVectorArray) -> na.ScalarArray:
        """``1 / self(coordinates)`` (``_vignetting.py:44-60``)."""
        return 1 / self(coordinates)


@dataclasses.dataclass(eq=False)
class PolynomialVignettingModel(AbstractVignettingModel):
    """
    Relative illumination as a polynomial of (wavelength, field) of total degree `degree` about the
    mean scene coordinate, fitted to the points selected by `where`
    (``optika/radiometry/_vignetting.py:99-180``).
    """

    coordinates_scene: SpectralPositionalVectorArray = None
    illumination: na.ScalarArray = None
    axis_wavelength: str = None
    axis_field: tuple = None
    degree: int = 1
    where: object = True

    @property
    def _axis_scene(self) -> tuple:
        return (self.axis_wavelength, *self.axis_field)

    @functools.cached_property
    def fit(self) -> PolynomialFit:
        scene = self.coordinates_scene
        inputs = (scene.wavelength, scene.position.x, scene.position.y)
        shape_ = na.broadcast_shapes(*[na.shape(a) for a in inputs], na.shape(self.illumination))
        inputs = tuple(na.broadcast_to(na.as_named_array(a), {ax: n for ax, n in shape_.items()
                                                              if ax in self._axis_scene or ax in na.shape(a)})
                       for a in inputs)
        return PolynomialFit(
            inputs=inputs,
            outputs=(self.illumination,),
            degree=self.degree,
            center=tuple(_mean(a, self._axis_scene) for a in inputs),  # _vignetting.py:168
            where=self.where,
            axes=tuple(ax for ax in shape_ if ax in self._axis_scene),
        )

    def __call__(self, coordinates: SpectralPositionalVectorArray) -> na.ScalarArray:
        (result,) = self.fit(coordinates.wavelength, coordinates.position.x, coordinates.position.y)
        return result

    @property
    def residual(self) -> na.ScalarArray:
        (prediction,) = self.fit.predictions
        r = na.as_named_array(self.illumination) - prediction
        where = na.broadcast_to(na.as_named_array(self.where), r.shape)
        return na.ScalarArray(np.where(where.ndarray, r.ndarray, np.nan), r.axes)


@dataclasses.dataclass(eq=False)
class AbstractEffectiveAreaModel:
    """``optika.radiometry.AbstractEffectiveAreaModel`` (``_effective_area.py:14-34``)."""

    def __call__(self, wavelength) -> na.ScalarArray:
        raise NotImplementedError


@dataclasses.dataclass(eq=False)
class InterpolatedEffectiveAreaModel(AbstractEffectiveAreaModel):
    """
    Effective area by linear interpolation in wavelength between calibration points
    (``optika/radiometry/_effective_area.py:37-102``: ``na.interp`` along `axis_wavelength`,
    i.e. ``numpy.interp`` with clamped ends, independently for every other axis of `area`).
    """

    wavelength: na.ScalarArray = None
    area: na.ScalarArray = None
    axis_wavelength: str = None

    def __call__(self, wavelength) -> na.ScalarArray:
        x = na.as_named_array(wavelength)
        xp = na.as_named_array(self.wavelength)
        fp = na.as_named_array(self.area)
        axis = self.axis_wavelength
        if xp.axes != (axis,):
            raise ValueError(f"the calibration wavelengths must vary along {axis!r} only, got {xp.axes}")
        order = np.argsort(xp.ndarray)
        others = {ax: n for ax, n in fp.shape.items() if ax != axis}
        full = dict(others, **{axis: fp.shape[axis]})
        values = np.broadcast_to(na.aligned(fp, full), tuple(full.values())).reshape(-1, fp.shape[axis])
        out_axes = tuple(others) + x.axes
        flat = np.stack([np.interp(x.ndarray.reshape(-1), xp.ndarray[order], row[order]) for row in values])
        return na.ScalarArray(flat.reshape(tuple(others.values()) + x.ndarray.shape), out_axes)
