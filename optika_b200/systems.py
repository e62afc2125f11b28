"""
Sequential optical systems.

Mirrors ``optika.systems.SequentialSystem`` (``optika/systems/_sequential.py``)
for the hot path: ``raytrace`` (``:836-925``) and ``rayfunction`` (``:927-988``)
build the input rays from the (wavelength, field, pupil) grid
(``_calc_rayfunction_input``, ``:791-828``) and hand the whole surface list to
the device engine in one call.  ``image_rays`` is the fused
trace-and-bin path used for detector image simulation
(``image`` -> ``sensor.measure``, ``:1088-1206``): rays never leave the chip.

Grid coordinates are PHYSICAL by default here (``normalized_field=False``):
the stop solver that maps normalised coordinates (``:396-678``) is a host-side
caller of the hot path and is provided separately by
:meth:`SequentialSystem.denormalize`.
"""

from __future__ import annotations
from typing import Sequence
import dataclasses
import functools
import numpy as np
from . import named as na
from . import units as u
from . import _util
from . import _engine
from . import _lib as L
from .rays import RayVectorArray, RayFunctionArray
from .surfaces import Surface, AbstractSurface
from .vectors import ObjectVectorArray
from .transformations import AbstractTransformation

__all__ = ["AbstractSequentialSystem", "SequentialSystem"]


@dataclasses.dataclass(eq=False)
class AbstractSequentialSystem:
    pass


@dataclasses.dataclass(eq=False)
class SequentialSystem(AbstractSequentialSystem):
    """A sequence of surfaces traced in order (``_sequential.py:1826-2132``)."""

    surfaces: Sequence[AbstractSurface] = ()
    object: None | AbstractSurface = None
    sensor: None | AbstractSurface = None
    grid_input: None | ObjectVectorArray = None
    axis_surface: str = "surface"
    transformation: None | AbstractTransformation = None
    object_at_infinity: None | bool = None
    """
    Stand-in for the unit test of ``object_is_at_infinity`` (``:44-62``): ``None``
    follows the reference rule (no object aperture, or an angular one => infinity).
    """

    def __post_init__(self):
        if self.object is None:
            self.object = Surface()  # _sequential.py:2121-2123

    # -- structure ---------------------------------------------------------
    @property
    def shape(self) -> dict[str, int]:
        # _sequential.py:2125-2132
        return na.broadcast_shapes(
            *[na.shape(s) for s in self.surfaces],
            na.shape(self.object),
            na.shape(self.sensor),
            na.shape(self.transformation),
        )

    @property
    def object_is_at_infinity(self) -> bool:
        # _sequential.py:44-62
        if self.object_at_infinity is not None:
            return self.object_at_infinity
        aperture = self.object.aperture
        if aperture is None:
            return True
        return bool(getattr(aperture, "angular", False))

    @property
    def surfaces_all(self) -> list[AbstractSurface]:
        # _sequential.py:93-111
        result = [self.object] if self.object is not None else []
        result += list(self.surfaces)
        if self.sensor is not None:
            result += [self.sensor]
        return result

    @functools.cached_property
    def _compiled(self) -> _engine.CompiledSystem:
        """The lowered surface table; cached like ``rayfunction_default`` (``:990-1000``)."""
        return _engine.CompiledSystem(self.surfaces_all)

    def invalidate(self):
        """Drop the cached device table after mutating a surface."""
        self.__dict__.pop("_compiled", None)

    # -- input rays --------------------------------------------------------
    def _calc_rayfunction_input(self, grid: ObjectVectorArray) -> RayFunctionArray:
        """Physical grid -> input rays, ``_sequential.py:791-828``."""
        if self.object_is_at_infinity:
            position = grid.pupil
            direction = _util.direction(grid.field)
        else:
            position = grid.field
            direction = _util.direction(grid.pupil)
        rays = RayVectorArray(
            wavelength=u.length(grid.wavelength),
            position=na.Cartesian3dVectorArray(
                x=u.length(position.x), y=u.length(position.y), z=0.0
            ),
            direction=direction,
        )
        obj = self.object
        if obj is not None and obj.transformation is not None:
            rays = obj.transformation(rays)
        return RayFunctionArray(inputs=grid, outputs=rays)

    def _input(self, intensity, wavelength, field, pupil, normalized_field, normalized_pupil):
        grid = self.grid_input.copy_shallow() if self.grid_input is not None else ObjectVectorArray()
        if wavelength is not None:
            grid.wavelength = wavelength
        if field is not None:
            grid.field = field
        if pupil is not None:
            grid.pupil = pupil
        if normalized_field or normalized_pupil:
            grid = self.denormalize(grid, normalized_field, normalized_pupil)
        result = self._calc_rayfunction_input(grid)
        rays = result.outputs
        if intensity is not None:
            rays.intensity = intensity
        if self.transformation is not None:
            rays = self.transformation.inverse(rays)  # _sequential.py:908-909
        # preferred device order of the ray axes: wavelength, field, pupil (pupil innermost)
        order = []
        for part in (grid.wavelength, grid.field, grid.pupil):
            for ax in na.shape(part):
                if ax not in order:
                    order.append(ax)
        self._ray_axes_order = order
        return result, rays

    def denormalize(self, grid: ObjectVectorArray, normalized_field=True, normalized_pupil=True, backend=None):
        """Map normalised field / pupil coordinates to physical ones (``:748-789``)."""
        from . import _stops

        return _stops.denormalize_grid(self, grid, normalized_field, normalized_pupil, backend=backend)

    def rayfunction_stops(self, wavelength=None, samples_pupil_stop=101, samples_field_stop=101, backend=None):
        """Rays through the edges of both stops, at the object (``:625-678``): ``(inputs, rays)``."""
        from . import _stops

        if wavelength is None:
            wavelength = self.grid_input.wavelength
        return _stops.rayfunction_stops(self, wavelength, samples_pupil_stop, samples_field_stop, backend=backend)

    def _stop_extent(self, backend=None):
        from . import _stops

        _, rays = self.rayfunction_stops(backend=backend)
        axes = (_stops.AXIS_FIELD_STOP, _stops.AXIS_PUPIL_STOP)
        if self.object_is_at_infinity:
            field = _util.angles(rays.direction)
            pupil = na.Cartesian2dVectorArray(rays.position.x, rays.position.y)
        else:
            field = na.Cartesian2dVectorArray(rays.position.x, rays.position.y)
            pupil = _util.angles(rays.direction)
        return field, pupil, axes

    def field_min(self, backend=None) -> na.Cartesian2dVectorArray:
        """Lower-left corner of the field of view (``_sequential.py:697-708``)."""
        field, _, axes = self._stop_extent(backend)
        return field.min(axes)

    def field_max(self, backend=None) -> na.Cartesian2dVectorArray:
        """Upper-right corner of the field of view (``_sequential.py:710-720``)."""
        field, _, axes = self._stop_extent(backend)
        return field.max(axes)

    def pupil_min(self, backend=None) -> na.Cartesian2dVectorArray:
        _, pupil, axes = self._stop_extent(backend)
        return pupil.min(axes)

    def pupil_max(self, backend=None) -> na.Cartesian2dVectorArray:
        """Upper-right corner of the entrance pupil (``_sequential.py:736-746``)."""
        _, pupil, axes = self._stop_extent(backend)
        return pupil.max(axes)

    # -- tracing -----------------------------------------------------------
    def raytrace(
        self,
        intensity=None,
        wavelength=None,
        field=None,
        pupil=None,
        axis: None | str = None,
        normalized_field: bool = False,
        normalized_pupil: bool = False,
        accumulate: bool = True,
        device=None,
        on_device: bool = False,
    ) -> RayFunctionArray:
        """
        Trace the input grid through the whole system; results in GLOBAL
        coordinates, optionally with the rays at every surface on `axis`
        (``_sequential.py:836-925``).
        """
        if axis is None:
            axis = self.axis_surface
        result, rays = self._input(intensity, wavelength, field, pupil, normalized_field, normalized_pupil)
        out = _engine.trace(
            self._compiled, rays, accumulate=accumulate, axis=axis, device=device,
            ray_axes_order=self._ray_axes_order,
        )
        result.outputs = out if on_device else out.to_host()
        return result

    def rayfunction(
        self,
        intensity=None,
        wavelength=None,
        field=None,
        pupil=None,
        normalized_field: bool = False,
        normalized_pupil: bool = False,
        device=None,
        on_device: bool = False,
    ) -> RayFunctionArray:
        """
        Rays at the last surface in the LOCAL coordinates of the sensor
        (``_sequential.py:927-988``).  Unlike the reference, which always pays
        for ``accumulate=True`` and then indexes the last surface (``:970-980``),
        only the final state is written.
        """
        result, rays = self._input(intensity, wavelength, field, pupil, normalized_field, normalized_pupil)
        # sensor.transformation.inverse(rays) == never leaving the sensor's local frame
        out = _engine.trace(self._compiled_local, rays, device=device, ray_axes_order=self._ray_axes_order)
        result.outputs = out if on_device else out.to_host()
        return result

    @functools.cached_property
    def _compiled_local(self) -> _engine.CompiledSystem:
        """
        The system with the final local->global step of the sensor removed, so the
        trace ends in sensor-local coordinates (``_sequential.py:983-986``).
        """
        return _engine.CompiledSystem(self.surfaces_all, local_last=True)

    def image_rays(
        self,
        wavelength_edges: na.ScalarArray,
        intensity=None,
        wavelength=None,
        field=None,
        pupil=None,
        normalized_field: bool = False,
        normalized_pupil: bool = False,
        device=None,
        counts: bool = True,
        image: None | _engine.DeviceImage = None,
    ) -> _engine.DeviceImage:
        """
        Fused trace + detector binning: the rays of the grid are traced to the
        sensor and binned into its pixel grid inside the same kernel launch
        (``image`` -> ``rayfunction`` -> ``sensor.collect``,
        ``_sequential.py:1077-1086, 1200-1206`` and ``sensors/_sensors.py:92-171``);
        no ray is ever written to HBM.  One image per configuration.
        Returns the device-resident planes (``flux``, ``moment_real``, ``counts``).
        """
        device = _engine.require_cuda(device)
        result, rays = self._input(intensity, wavelength, field, pupil, normalized_field, normalized_pupil)
        ex, ey = self.sensor.pixel_edges()
        edges_w = np.asarray(na.as_named_array(u.length(wavelength_edges)).ndarray, dtype=float)
        # trace in the sensor's local frame straight away (no final local -> global step, no
        # inverse frame transformation before binning)
        compiled = self._compiled_local
        if image is None:
            image = _engine.DeviceImage.zeros(
                edges_w, ex, ey, device, leading=tuple(compiled.shape.values()), moments=True, counts=counts
            )
        _engine.trace(
            compiled, rays, image=image, image_frame=None, write_rays=False, device=device,
            ray_axes_order=self._ray_axes_order,
        )
        return image
