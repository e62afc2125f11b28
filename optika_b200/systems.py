"""
Sequential optical systems.

Mirrors ``optika.systems.SequentialSystem`` (``optika/systems/_sequential.py``)
for the hot path: ``raytrace`` (``:836-925``) and ``rayfunction`` (``:927-988``)
build the input rays from the (wavelength, field, pupil) grid
(``_calc_rayfunction_input``, ``:791-828``) and hand the whole surface list to
the device engine in one call.  ``image_rays`` is the fused
trace-and-bin path used for detector image simulation
(``image`` -> ``sensor.measure``, ``:1088-1206``): rays never leave the chip.

Grid coordinates are NORMALISED by default, as in the reference (``:843-844, 936-937``): they
are mapped to physical ones through the stop solver (``:396-678``, on the device:
``optk_solve_stops``), solved once per wavelength grid and cached.  Floats carry no unit here, so a
grid in degrees / millimetres must be announced with ``normalized_field=False,
normalized_pupil=False`` (the reference would raise a unit error instead).
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
    coating: str = "exact"
    """
    How multilayer-coated surfaces are evaluated: ``"exact"`` -- ``multilayer_efficiency`` for every ray, as the
    reference does (``optika/materials/_multilayers.py:908-935``; the trace is chained through HBM around
    the coated surface); ``"table"`` -- efficiency(wavelength, cosine of incidence) tabulated once to
    `coating_tolerance` and looked up inside one fused launch (:mod:`optika_b200._coatings`).
    """
    coating_tolerance: float = 1e-6
    """Largest interpolation error of a coating table, relative to the largest efficiency in it."""
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

    def _compile(self, local_last: bool) -> _engine.CompiledSystem:
        """
        The surface list lowered to the device table.  The reference re-reads ``surfaces_all`` on every
        ``raytrace`` (``_sequential.py:913-923``), so the list is lowered on every call (cheap host work)
        and the device handle of the previous call is reused only when the lowered bytes are the same:
        editing a radius, a transformation or the sensor takes effect on the next trace.
        """
        from . import _lowering

        surfaces = self.surfaces_all
        cache = self.__dict__.setdefault("_compiled_cache", {})
        entry = cache.get(local_last)
        # cheap change detection first: a digest of every field of every surface (0.2 ms against the
        # 2-20 ms of lowering a system with a configuration axis)
        digest = _lowering.fingerprint(surfaces)
        if entry is None or entry[2] != digest:
            lowered = _lowering.lower_system(surfaces)
            key = _lowering.table_key(lowered[0])
            if entry is None or entry[0] != key or len(entry[1].surfaces) != len(surfaces):
                entry = (key, _engine.CompiledSystem(surfaces, local_last=local_last, lowered=lowered), digest)
            else:
                entry = (entry[0], entry[1], digest)
            cache[local_last] = entry
        compiled = entry[1]
        compiled.coating, compiled.coating_tolerance = self.coating, self.coating_tolerance
        # objects that are read again at trace time (coating stacks) always come from the current list
        compiled.surfaces = surfaces
        compiled.coatings = {k: s.material for k, s in enumerate(surfaces) if hasattr(s.material, "efficiency_device")}
        return compiled

    @property
    def _compiled(self) -> _engine.CompiledSystem:
        return self._compile(local_last=False)

    def invalidate(self):
        """Kept for callers of round 1: the device tables now follow the surfaces by themselves."""
        self.__dict__.pop("_compiled_cache", None)
        self.__dict__.pop("_stop_cache", None)

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
        from . import _stops, _lowering

        if (not normalized_field) and (not normalized_pupil):
            return grid
        # The stop solution depends on the surfaces and on the wavelength grid only: solved once and
        # kept (the reference caches it for its default grid, ``_rayfunction_input``, :830-834), keyed
        # on the lowered surface table so that an edited surface is solved again.
        w = na.as_named_array(u.length(grid.wavelength))
        frame = self.transformation.affine.numpy({}) if self.transformation is not None and not na.shape(self.transformation) else None
        key = (
            _lowering.table_key(_lowering.lower_system(self.surfaces_all)[0]),
            tuple(w.axes), np.ascontiguousarray(w.ndarray, dtype=float).tobytes(),
            None if frame is None else (frame[0].tobytes(), frame[1].tobytes()),
            self.object_is_at_infinity, backend,
        )
        cache = self.__dict__.setdefault("_stop_cache", {})
        cacheable = self.transformation is None or frame is not None
        extents = cache.get(key) if cacheable else None
        if extents is None:
            extents = _stops.stop_extents(self, grid.wavelength, backend)
            if cacheable:
                if len(cache) >= 8:
                    cache.pop(next(iter(cache)))
                cache[key] = extents
        return _stops.denormalize_grid(self, grid, normalized_field, normalized_pupil, backend=backend, extents=extents)

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
        normalized_field: bool = True,
        normalized_pupil: bool = True,
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
        normalized_field: bool = True,
        normalized_pupil: bool = True,
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

    def pupil_moments(
        self,
        intensity=None,
        wavelength=None,
        field=None,
        pupil=None,
        normalized_field: bool = True,
        normalized_pupil: bool = True,
        device=None,
        axis_pupil: None | tuple = None,
    ) -> dict:
        """
        What ``distortion``, ``vignetting`` and ``area_effective`` reduce from the rays at the
        sensor (``_sequential.py:1266-1285, 1351-1368, 1501-1506``), computed on the device: the
        rays are traced to the sensor (local coordinates, as ``rayfunction``) and reduced over the
        pupil axes for every configuration, wavelength and field point inside the same kernel launch
        (``optk_image_t.group_size``): no ray is written to HBM, so the grid may be as large as an
        ``image`` grid.

        Returns named arrays over the remaining axes:

        * ``where``: ``unvignetted.any(axis_pupil)``;
        * ``position``: ``mean(position.xy, axis_pupil, where=unvignetted | ~where)`` -- the mean
          over the unvignetted rays, or over all rays where none survives (``:1272-1279``);
        * ``illumination``: ``unvignetted.mean(axis_pupil)`` (``:1354``, before its normalisation);
        * ``intensity``: ``intensity.sum(axis_pupil, where=unvignetted)`` (``:1501-1504``).
        """
        device = _engine.require_cuda(device)
        result, rays = self._input(intensity, wavelength, field, pupil, normalized_field, normalized_pupil)
        # the axes reduced over: those of the pupil grid, or the ones the caller names (a pupil sampled
        # independently for every wavelength and field point carries all the axes of the grid)
        pupil_axes = na.shape(result.inputs.pupil) if axis_pupil is None else axis_pupil
        axis_pupil = [ax for ax in self._ray_axes_order if ax in pupil_axes]
        compiled = self._compiled_local
        config, ray = _engine._grid_shape(rays, compiled.shape)
        missing = [ax for ax in pupil_axes if ax not in ray]
        if missing or not axis_pupil:
            raise ValueError(f"the pupil axes {missing} are not axes of the traced rays")
        # device order of the ray axes (see _engine._trace): the requested order last, pupil innermost
        order = [ax for ax in ray if ax not in self._ray_axes_order] + [ax for ax in self._ray_axes_order if ax in ray]
        if order[len(order) - len(axis_pupil):] != axis_pupil:
            raise ValueError("the pupil axes must be the innermost ray axes")
        outer = {ax: ray[ax] for ax in order[: len(order) - len(axis_pupil)]}
        n_inner = int(np.prod([ray[ax] for ax in axis_pupil], dtype=np.int64))
        n_groups = int(np.prod(list(outer.values()), dtype=np.int64)) if outer else 1
        n_config = int(np.prod(list(config.values()), dtype=np.int64)) if config else 1
        shape_ = dict(config)
        shape_.update(outer)
        axes, dims = tuple(shape_), tuple(shape_.values())
        fused = not compiled.coatings
        if fused:
            # the kernel that traces the rays also reduces them; no ray is written to HBM
            groups = _engine.DeviceGroups.zeros(n_config, n_groups, n_inner, device)
            try:
                _engine.trace(
                    compiled, rays, image=groups, write_rays=False, device=device, ray_axes_order=self._ray_axes_order
                )
            except NotImplementedError:
                fused = False  # no run-time compiled kernel for this launch (generic operator, OPTK_JIT=0, no NVRTC)
        if not fused:
            # multilayer-coated surfaces are traced in chained launches with the rays in HBM between the
            # links, and launches without a run-time compiled kernel: reduce the dense result
            out = _engine.trace(compiled, rays, device=device, ray_axes_order=self._ray_axes_order)
            sums = _engine.reduce_groups(out, axis_pupil, device=device)
            count, sum_i = sums["count"].ndarray.reshape(dims), sums["sum_intensity"].ndarray.reshape(dims)
            sum_x, sum_y = sums["sum_x"].ndarray.reshape(dims), sums["sum_y"].ndarray.reshape(dims)
            all_x, all_y = sums["sum_x_all"].ndarray.reshape(dims), sums["sum_y_all"].ndarray.reshape(dims)
        else:
            count = groups.counts.cpu().numpy().reshape(dims)
            sum_i = groups.flux.cpu().numpy().reshape(dims)
            sum_x = groups.moment_real.cpu().numpy().reshape(dims)
            sum_y = groups.moment_imag.cpu().numpy().reshape(dims)
            all_x = all_y = None
        where = count > 0
        if not where.all() and all_x is None:
            # field points without a surviving ray: the reference averages over ALL their rays there
            # (:1272-1279, a placeholder that its fits mask out); rare, so taken from a plain trace
            out = _engine.trace(compiled, rays, device=device, ray_axes_order=self._ray_axes_order)
            sums = _engine.reduce_groups(out, axis_pupil, device=device)
            all_x, all_y = sums["sum_x_all"].ndarray.reshape(dims), sums["sum_y_all"].ndarray.reshape(dims)
        denominator = np.where(where, count, n_inner)
        x = np.where(where, sum_x, all_x if all_x is not None else 0.0) / denominator
        y = np.where(where, sum_y, all_y if all_y is not None else 0.0) / denominator
        return dict(
            inputs=result.inputs,
            where=na.ScalarArray(where, axes),
            position=na.Cartesian2dVectorArray(na.ScalarArray(x, axes), na.ScalarArray(y, axes)),
            illumination=na.ScalarArray(count / n_inner, axes),
            intensity=na.ScalarArray(sum_i, axes),
        )

    # -- consumers of ray output ----------------------------------
    def _scene_axes(self, grid: ObjectVectorArray, what: str):
        """``(axis_wavelength, axis_field)`` of a grid, as ``_rayfunction_and_axes`` finds them (``:1514-1582``)."""
        config_axes = set(self.shape)
        axis_wavelength = tuple(ax for ax in na.shape(grid.wavelength) if ax not in config_axes)
        if len(axis_wavelength) != 1:
            raise ValueError(
                f"fitting a {what} model requires the wavelength grid to vary along its own logical axis."
            )  # :1254-1258, :1339-1343
        skip = config_axes | set(axis_wavelength)
        axis_field = tuple(ax for ax in na.shape(grid.field) if ax not in skip)
        return axis_wavelength[0], axis_field

    def distortion(
        self,
        wavelength=None,
        field=None,
        pupil=None,
        normalized_field: bool = True,
        normalized_pupil: bool = True,
        degree: int = 2,
        device=None,
    ):
        """
        Fit a polynomial distortion model to the rays traced through this system
        (``optika/systems/_sequential.py:1208-1285``).  The per-field-point mean sensor position over
        the unvignetted rays of the pupil (over all of them where none survives, ``:1269-1276``) comes
        out of the trace kernel itself (:meth:`pupil_moments`: no ray is written to HBM); the
        least-squares fit of the few hundred field points is host work.
        """
        from .distortion import PolynomialDistortionModel
        from .vectors import SpectralPositionalVectorArray

        moments = self.pupil_moments(None, wavelength, field, pupil, normalized_field, normalized_pupil, device=device)
        grid = moments["inputs"]
        axis_wavelength, axis_field = self._scene_axes(grid, "distortion")
        return PolynomialDistortionModel(
            coordinates_scene=SpectralPositionalVectorArray(wavelength=grid.wavelength, position=grid.field),
            coordinates_sensor=moments["position"],
            axis_wavelength=axis_wavelength,
            axis_field=axis_field,
            degree=degree,
            where=moments["where"],
        )

    def vignetting(
        self,
        wavelength=None,
        field=None,
        pupil=None,
        normalized_field: bool = True,
        normalized_pupil: bool = True,
        degree: int = 2,
        device=None,
    ):
        """
        Fit a polynomial vignetting model (``optika/systems/_sequential.py:1287-1368``): the relative
        illumination of a scene coordinate is the unvignetted fraction of its pupil, normalised to a
        unit mean over the field points that keep at least one ray (``:1351-1358``).
        """
        from .radiometry import PolynomialVignettingModel
        from .vectors import SpectralPositionalVectorArray

        moments = self.pupil_moments(None, wavelength, field, pupil, normalized_field, normalized_pupil, device=device)
        grid = moments["inputs"]
        axis_wavelength, axis_field = self._scene_axes(grid, "vignetting")
        where, illumination = moments["where"], moments["illumination"]
        axes = illumination.axes
        keep = tuple(k for k, ax in enumerate(axes) if ax in axis_field)
        w = where.ndarray
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = (illumination.ndarray * w).sum(axis=keep, keepdims=True) / w.sum(axis=keep, keepdims=True)
            illumination = na.ScalarArray(illumination.ndarray / mean, axes)
        return PolynomialVignettingModel(
            coordinates_scene=SpectralPositionalVectorArray(wavelength=grid.wavelength, position=grid.field),
            illumination=illumination,
            axis_wavelength=axis_wavelength,
            axis_field=axis_field,
            degree=degree,
            where=where,
        )

    def area_effective(
        self,
        wavelength=None,
        field=None,
        pupil=None,
        normalized_field: bool = True,
        normalized_pupil: bool = True,
        seed: int = 0,
        device=None,
    ):
        """
        The wavelength-dependent effective area (``optika/systems/_sequential.py:1370-1512``): `pupil`
        holds the VERTICES of a grid of pupil cells (default 11 x 11 over the normalised pupil, ``:1440-1444``);
        one stratified random ray per cell, wavelength and field point carries the area of its cell as
        intensity (``:1490-1499``), the traced intensities of the unvignetted rays are summed over the
        pupil (inside the trace kernel) and averaged over the field (``:1501-1506``).  `seed` selects
        the jitter (the reference draws from NumPy's global generator).
        """
        from .radiometry import InterpolatedEffectiveAreaModel

        if wavelength is None:
            wavelength = self.grid_input.wavelength
        if field is None:
            field = self.grid_input.field
        if pupil is None:
            pupil = na.Cartesian2dVectorArray(
                x=na.linspace(-1, 1, axis="_pupil_x", num=11), y=na.linspace(-1, 1, axis="_pupil_y", num=11)
            )
        grid = self.denormalize(ObjectVectorArray(wavelength=wavelength, field=field, pupil=pupil),
                                normalized_field, normalized_pupil)
        wavelength, field, pupil = grid.wavelength, grid.field, grid.pupil
        config_axes = set(self.shape)
        axis_wavelength = tuple(ax for ax in na.shape(wavelength) if ax not in config_axes)
        if len(axis_wavelength) != 1:
            raise ValueError(
                f"Computing the effective area requires that there be only one wavelength axis, got {axis_wavelength}"
            )  # :1477-1481
        (axis_wavelength,) = axis_wavelength
        skip = config_axes | {axis_wavelength}
        axis_field = tuple(ax for ax in na.shape(field) if ax not in skip)
        skip |= set(axis_field)
        axis_pupil = tuple(ax for ax in na.shape(pupil) if ax not in skip)
        if len(axis_pupil) != 2:
            raise ValueError(f"the pupil vertices must vary along two axes of their own, got {axis_pupil}")
        px, py = na.as_named_array(pupil.x), na.as_named_array(pupil.y)
        area = np.abs(na.Cartesian2dVectorArray(
            na.broadcast_to(px, na.broadcast_shapes(px.shape, py.shape)),
            na.broadcast_to(py, na.broadcast_shapes(px.shape, py.shape)),
        ).volume_cell(axis_pupil))  # :1485
        # cell_centers(axis_pupil, random=True) over the FULL grid shape (:1486): one uniform sample per
        # cell for every wavelength and field point
        full = na.broadcast_shapes(na.shape(wavelength), na.shape(field), na.shape(pupil))
        rng = np.random.default_rng(seed)
        axes_full = tuple(full)
        ka, kb = axes_full.index(axis_pupil[0]), axes_full.index(axis_pupil[1])
        cells = tuple(n - 1 if ax in axis_pupil else n for ax, n in full.items())
        ta, tb = rng.uniform(size=cells), rng.uniform(size=cells)
        centres = []
        for c in (pupil.x, pupil.y):
            nd = na.broadcast_to(na.as_named_array(c), full).ndarray

            def corner(da, db):
                index = [slice(None)] * nd.ndim
                index[ka] = slice(da, nd.shape[ka] - 1 + da)
                index[kb] = slice(db, nd.shape[kb] - 1 + db)
                return nd[tuple(index)]

            lo = corner(0, 0) + ta * (corner(1, 0) - corner(0, 0))  # along the first pupil axis, then the second
            hi = corner(0, 1) + ta * (corner(1, 1) - corner(0, 1))
            centres.append(na.ScalarArray(lo + tb * (hi - lo), axes_full))
        moments = self.pupil_moments(
            area, wavelength, field, na.Cartesian2dVectorArray(*centres), False, False, device=device,
            axis_pupil=axis_pupil,
        )
        area_eff = moments["intensity"]
        keep = tuple(ax for ax in area_eff.axes if ax in axis_field)
        area_eff = area_eff.mean(keep) if keep else area_eff  # :1506
        return InterpolatedEffectiveAreaModel(wavelength=wavelength, area=area_eff, axis_wavelength=axis_wavelength)

    @property
    def _compiled_local(self) -> _engine.CompiledSystem:
        """
        The system with the final local->global step of the sensor removed, so the
        trace ends in sensor-local coordinates (``_sequential.py:983-986``).
        """
        return self._compile(local_last=True)

    def image_rays(
        self,
        wavelength_edges: na.ScalarArray,
        intensity=None,
        wavelength=None,
        field=None,
        pupil=None,
        normalized_field: bool = True,
        normalized_pupil: bool = True,
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

    # -- forward model: scene -> detector image ------------------------------
    def _separable(self, value, axis: str, config_shape: dict, cindex: tuple, what: str,
                   axis_wavelength: None | str = None) -> np.ndarray:
        """
        The 1-D vertex array of a grid component that varies along `axis` (and possibly the
        configuration axes).  Other axes (e.g. the wavelength axis the stop solution is
        broadcast over, ``_sequential.py:760-789``) are accepted when the values are constant
        along them to 1e-9 of the extent of the grid.  Vertices that DO depend on the wavelength
        (chromatic stop solutions: the field of view of a spectrograph) come back as a 2-D array
        ``[n_wavelength + 1][n + 1]``, one row per wavelength vertex (``optk_grid_t.chromatic``).
        """
        v = na.as_named_array(value)
        if axis not in v.axes:
            raise ValueError(f"the {what} vertices must vary along axis {axis!r}, got axes {v.axes}")
        shape_ = dict(config_shape)
        for ax, n in v.shape.items():
            if ax not in shape_:
                shape_[ax] = n
        nd = np.broadcast_to(na.aligned(v, shape_), tuple(shape_.values()))[cindex]
        axes = [ax for ax in shape_ if ax not in config_shape]
        full = nd
        nd = np.moveaxis(nd, axes.index(axis), -1).reshape(-1, v.shape[axis])
        extent = float(np.ptp(nd)) or 1.0
        if float(np.ptp(nd, axis=0).max()) > 1e-9 * extent:
            if axis_wavelength is not None and axis_wavelength in axes and axis_wavelength != axis:
                # constant along everything but the wavelength axis?
                rows = np.moveaxis(full, [axes.index(axis_wavelength), axes.index(axis)], [-2, -1])
                rows = rows.reshape((-1,) + rows.shape[-2:])
                if float(np.ptp(rows, axis=0).max()) <= 1e-9 * extent:
                    return np.ascontiguousarray(rows.mean(axis=0), dtype=np.float64)
            raise NotImplementedError(
                f"the {what} vertices vary along axes other than {axis!r}: only separable grids run on the "
                "device generator; trace explicit rays with `image_rays` instead"
            )
        return np.asarray(nd.mean(axis=0), dtype=np.float64)

    def _curvilinear(self, vector, axes: tuple, config_shape: dict, cindex: tuple, convert, what: str):
        """
        The two 2-D vertex arrays ``[n_a + 1][n_b + 1]`` of a field / pupil grid whose components
        both vary along both of its axes (e.g. a polar pupil grid); other axes as in `_separable`.
        """
        out = []
        for c in (vector.x, vector.y):
            v = na.as_named_array(convert(c))
            shape_ = dict(config_shape)
            for ax in axes:
                n = na.shape(vector).get(ax)
                if n is None:
                    raise ValueError(f"the {what} vertices must vary along axes {axes}, got {na.shape(vector)}")
                shape_[ax] = n
            for ax, n in v.shape.items():
                shape_.setdefault(ax, n)
            nd = np.broadcast_to(na.aligned(v, shape_), tuple(shape_.values()))[cindex]
            names = [ax for ax in shape_ if ax not in config_shape]
            nd = np.moveaxis(nd, [names.index(axes[0]), names.index(axes[1])], [-2, -1])
            nd = nd.reshape((-1,) + nd.shape[-2:])
            extent = float(np.ptp(nd)) or 1.0
            if float(np.ptp(nd, axis=0).max()) > 1e-9 * extent:
                raise NotImplementedError(
                    f"the {what} vertices vary along axes other than {axes}; trace explicit rays with `image_rays`"
                )
            out.append(np.ascontiguousarray(nd.mean(axis=0), dtype=np.float64))
        return tuple(out)

    def _grid_vertices(self, vector, axes: tuple, config_shape: dict, cindex: tuple, convert, what: str,
                       axis_wavelength: None | str = None):
        """1-D (separable) vertex arrays of a field / pupil grid when possible -- 2-D per wavelength vertex
        when they depend on the wavelength -- else 2-D (curvilinear) ones."""
        x, y = na.as_named_array(convert(vector.x)), na.as_named_array(convert(vector.y))
        if axes[1] not in x.axes and axes[0] not in y.axes:
            return (
                self._separable(x, axes[0], config_shape, cindex, what + " x", axis_wavelength),
                self._separable(y, axes[1], config_shape, cindex, what + " y", axis_wavelength),
            )
        return self._curvilinear(vector, axes, config_shape, cindex, convert, what)

    def _frame_input(self, config_shape: dict, cindex: tuple):
        """Object-local -> first-surface coordinates as ``(R, t)`` (``_sequential.py:823-826, 908-909``)."""
        t_obj = self.object.transformation if self.object is not None else None
        if t_obj is None and self.transformation is None:
            return None
        affine = None
        if t_obj is not None:
            affine = t_obj.affine
        if self.transformation is not None:
            inv = self.transformation.affine.inverse
            affine = inv if affine is None else inv @ affine
        if not set(affine.shape).issubset(config_shape):
            raise NotImplementedError("system transformations with their own named axes are not supported here")
        r, t = affine.numpy(config_shape)
        return (r[cindex], t[cindex]) if cindex else (r, t)

    def ray_grids(
        self,
        radiance,
        wavelength,
        field,
        pupil,
        axis_wavelength: str,
        axis_field: tuple[str, str],
        axis_pupil: tuple[str, str],
        normalized_field: bool = True,
        normalized_pupil: bool = True,
        random: bool = True,
        seed: int = 0,
    ) -> list:
        """
        ``_rayfunction_from_vertices`` (``_sequential.py:1002-1086``) for separable grids:
        one :class:`~optika_b200._grid.RayGrid` per configuration, carrying the physical
        vertices, ``flux = radiance * cell_area`` split into its scene and pupil factors,
        and the object frame.  The rays themselves are only ever created on the device.
        """
        from ._grid import RayGrid

        grid = ObjectVectorArray(wavelength=wavelength, field=field, pupil=pupil)
        grid = self.denormalize(grid, normalized_field, normalized_pupil)
        config_shape = self._compiled_local.shape
        at_infinity = self.object_is_at_infinity
        conv_field, conv_pupil = (u.angle, u.length) if at_infinity else (u.length, u.angle)
        axes = (axis_wavelength,) + tuple(axis_field) + tuple(axis_pupil)
        # Vertices and weights that do not depend on the configuration (the usual case: a tolerance sweep
        # moves surfaces, not the scene) are computed and uploaded once and shared by every configuration.
        parts = (grid.wavelength, grid.field.x, grid.field.y, grid.pupil.x, grid.pupil.y, radiance)
        shared = not any(set(na.shape(part)) & set(config_shape) for part in parts)
        grids = []
        base = None
        for cindex in np.ndindex(*config_shape.values()) if config_shape else [()]:
            if shared and base is not None:
                g = dataclasses.replace(base, frame=self._frame_input(config_shape, cindex))
                g._device = base._device
                grids.append(g)
                continue
            shape_c, index_c = ({}, ()) if shared else (config_shape, cindex)
            vertices = (
                self._separable(u.length(grid.wavelength), axis_wavelength, shape_c, index_c, "wavelength"),
                *self._grid_vertices(grid.field, axis_field, shape_c, index_c, conv_field, "field", axis_wavelength),
                *self._grid_vertices(grid.pupil, axis_pupil, shape_c, index_c, conv_pupil, "pupil", axis_wavelength),
            )
            # axes whose vertices depend on the wavelength: separable in their pair (each component varies
            # along its own axis only), one row per wavelength vertex
            def separable(vector, pair):
                return pair[1] not in na.shape(vector.x) and pair[0] not in na.shape(vector.y)

            chromatic = tuple(
                a for a in (1, 2, 3, 4)
                if vertices[a].ndim == 2 and separable(grid.field if a < 3 else grid.pupil, axis_field if a < 3 else axis_pupil)
            )

            def named(a, b, axes2):
                if a in chromatic or b in chromatic:
                    one = lambda c, ax: na.ScalarArray(vertices[c], (axis_wavelength, ax) if c in chromatic else ax)  # noqa: E731
                    return na.Cartesian2dVectorArray(one(a, axes2[0]), one(b, axes2[1]))
                if vertices[a].ndim == 2:
                    return na.Cartesian2dVectorArray(na.ScalarArray(vertices[a], axes2), na.ScalarArray(vertices[b], axes2))
                return na.Cartesian2dVectorArray(
                    na.ScalarArray(vertices[a], axes2[0]), na.ScalarArray(vertices[b], axes2[1])
                )

            cell = ObjectVectorArray(
                wavelength=na.ScalarArray(vertices[0], axis_wavelength),
                field=named(1, 2, tuple(axis_field)),
                pupil=named(3, 4, tuple(axis_pupil)),
            )
            area_w, area_f, area_p = cell.cell_area(
                axis_wavelength, axis_field, axis_pupil,
                field_is_angular=at_infinity, pupil_is_angular=not at_infinity, factors=True,
            )
            if 1 in chromatic or 2 in chromatic:
                n_field = [vertices[c].shape[-1] - 1 for c in (1, 2)]
            else:
                n_field = [s_ - 1 for s_ in vertices[1].shape] if vertices[1].ndim == 2 else [len(vertices[1]) - 1, len(vertices[2]) - 1]
            n = [len(vertices[0]) - 1] + n_field
            scene_shape = dict(zip(axes[:3], n[:3]))
            rad = na.as_named_array(radiance)
            extra = set(rad.axes) - set(scene_shape) - set(config_shape)
            if extra:
                raise ValueError(f"the radiance has axes {sorted(extra)} that are not scene or configuration axes")
            full = dict(shape_c, **scene_shape)
            rad = np.broadcast_to(na.aligned(rad, full), tuple(full.values()))[index_c]
            # (with chromatic vertices the cell areas carry the wavelength axis: one value per wavelength cell)
            weight_scene = rad * area_w.numpy(axes[:1])[:, None, None] * na.as_named_array(area_f).numpy(axes[:3])
            weight_pupil = na.as_named_array(area_p).numpy((axes[0],) + tuple(axes[3:]))
            if weight_pupil.shape[0] == 1:
                weight_pupil = weight_pupil[0]
            base = RayGrid(
                vertices=vertices,
                chromatic=chromatic,
                at_infinity=at_infinity,
                weight_scene=weight_scene,
                weight_pupil=weight_pupil,
                jitter=random,
                seed=seed,
                frame=self._frame_input(config_shape, cindex),
                axes=axes,
            )
            grids.append(base)
        return grids

    def collect_grids(self, grids: list, wavelength_edges, device=None, reduce: bool = True, counts: bool = False,
                      pipeline=None, image=None, shard: bool = True, on_launch=None, moments: bool = True) -> dict:
        """
        The device part of :meth:`image`: every ray of `grids` (one :class:`~optika_b200._grid.RayGrid`
        per configuration, from :meth:`ray_grids`) is drawn, traced and binned on the device
        (``optk_trace_grid``: fused, no ray touches HBM) and the detector planes come back as host
        arrays ``{"flux"[, "moment_real"][, "counts"]}`` of shape ``[config axes..., n_w, n_x, n_y]``
        (`moments`: the flux x cos(incidence) plane behind ``sensor.collect``'s direction, ``sensors/_sensors.py:163-169``).

        Under ``torch.distributed`` every rank traces a slab of the grid (the axis that balances best,
        :func:`optika_b200.distributed.best_shard_axis`); while configuration k + 1 is traced, the planes
        of configuration k are summed over the ranks (``reduce_scatter``, NCCL) and each rank copies its
        slice to a shared page-locked host buffer over its own PCIe link
        (:class:`optika_b200.distributed.ImagePipeline`).  `pipeline` / `image` let a caller that
        simulates many exposures keep the buffers; ``shard=False`` traces the whole grid in this process
        (with a ``local`` pipeline: no collective); `on_launch(c, phase)` is called before (``0``) and after
        (``1``) the launches of configuration `c` are queued (the bench records CUDA events there).
        """
        from . import _grid, distributed

        device = _engine.require_cuda(device)
        compiled = self._compiled_local
        ex, ey = self.sensor.pixel_edges()
        rank, world = distributed.rank_world() if shard else (0, 1)
        own = pipeline is None
        with _engine.device_guard(device):
            if image is None:
                image = _engine.DeviceImage.zeros(
                    np.asarray(wavelength_edges, dtype=float), ex, ey, device, leading=tuple(compiled.shape.values()),
                    moments=moments, counts=counts, fused=True, pad_to=world if reduce else 1,
                ) if pipeline is None else pipeline.image
            if pipeline is None:
                pipeline = distributed.ImagePipeline(image, device, local=(world == 1)) if (reduce or world == 1) else None
            try:
                for c, grid in enumerate(grids):
                    if world > 1:
                        grid = grid.shard(rank, world, axis=distributed.best_shard_axis(grid.count, world))
                    if on_launch is not None:
                        on_launch(c, 0)
                    if compiled.coatings:
                        _grid.trace_grid_coated(compiled, grid, c, image, device=device)
                    else:
                        _grid.trace_grid(compiled, grid, config=c, image=image, write_rays=False, device=device)
                    if on_launch is not None:
                        on_launch(c, 1)
                    if pipeline is not None:
                        pipeline.submit(c)
                if pipeline is None:
                    return {k: v.numpy() for k, v in image.to_host(pinned=False).items()}
                planes = pipeline.finish()
                for tabled in compiled.__dict__.pop("_tabled_pending", []):
                    tabled.check()  # rays outside an efficiency table: fail loudly, not with a clamped value
                return {k: np.array(v) for k, v in planes.items()} if own else planes
            finally:
                if own and pipeline is not None:
                    pipeline.close()

    def image(
        self,
        scene: na.FunctionArray,
        pupil: None | na.Cartesian2dVectorArray = None,
        axis_wavelength: None | str = None,
        axis_field: None | tuple[str, str] = None,
        axis_pupil: None | tuple[str, str] = None,
        integrate: bool = True,
        noise: bool = True,
        normalized_field: bool = False,
        normalized_pupil: bool = True,
        seed: int = 0,
        device=None,
        reduce: bool = True,
    ) -> na.FunctionArray:
        """
        Forward model: spectral radiance of a scene -> detector counts
        (``optika/systems/_sequential.py:1088-1206``).

        `scene.inputs` holds the cell VERTICES of the wavelength and field grids and
        `scene.outputs` the radiance of every cell, per mm of wavelength, per sr (or mm^2)
        of field and per mm^2 (or sr) of pupil; `pupil` the vertices of the pupil grid
        (default: one cell spanning the whole normalised pupil, ``:1139-1146``).  One
        stratified random ray per cell is drawn, traced and binned on the device in a single
        fused launch per configuration (``optk_trace_grid``); ``sensor.expose`` then turns the
        photon image into electrons (``:1200-1206``, ``sensors/_sensors.py:374-428``).

        Differences from the reference signature: floats carry no unit, so
        `normalized_field` / `normalized_pupil` say whether the vertices are normalised
        (the reference infers it from the unit, ``:1148-1152``); `seed` selects the
        counter-based random stream.  Under ``torch.distributed`` every rank traces a
        slab of the pupil (or field) cells and the planes are summed (`reduce`).
        """
        from . import _grid, distributed

        device = _engine.require_cuda(device)
        wavelength = scene.inputs.wavelength
        field = scene.inputs.position
        if pupil is None:
            axis_pupil = ("_pupil_x", "_pupil_y")
            pupil = na.Cartesian2dVectorLinearSpace(-1, 1, na.Cartesian2dVectorArray(*axis_pupil), 2)
            normalized_pupil = True
        config_axes = set(self.shape)
        if axis_wavelength is None:
            cand = tuple(set(na.shape(wavelength)) - config_axes)
            if len(cand) != 1:
                raise ValueError(f"`scene` must vary along exactly one wavelength axis, got {cand} as possibilities.")
            (axis_wavelength,) = cand
        if axis_field is None:
            fx = set(na.shape(field.x)) - config_axes - {axis_wavelength}
            fy = set(na.shape(field.y)) - config_axes - {axis_wavelength}
            if len(fx) != 1 or len(fy) != 1 or fx == fy:
                raise ValueError(f"the two field axes must be unambiguous, got {sorted(fx | fy)} as possibilities.")
            axis_field = (next(iter(fx)), next(iter(fy)))
        if axis_pupil is None:
            skip = config_axes | {axis_wavelength} | set(axis_field)
            px, py = set(na.shape(pupil.x)) - skip, set(na.shape(pupil.y)) - skip
            if len(px) != 1 or len(py) != 1 or px == py:
                raise ValueError(f"the two pupil axes must be unambiguous, got {sorted(px | py)} as possibilities.")
            axis_pupil = (next(iter(px)), next(iter(py)))

        grids = self.ray_grids(
            scene.outputs, wavelength, field, pupil, axis_wavelength, tuple(axis_field), tuple(axis_pupil),
            normalized_field, normalized_pupil, random=True, seed=seed,
        )
        w_edges = np.asarray(na.as_named_array(u.length(wavelength)).ndarray, dtype=float)
        if na.as_named_array(wavelength).ndim != 1:
            raise NotImplementedError("the wavelength vertices of the scene must be one-dimensional")
        if integrate:
            w_edges = np.array([w_edges.min(), w_edges.max()])  # :1189-1196
        compiled = self._compiled_local
        ex, ey = self.sensor.pixel_edges()
        sensor = self.sensor
        # the angle of incidence on the detector is binned (and reduced, and read back) only for sensor materials
        # that look at it; the default IdealSensorMaterial does not (sensors/materials/_materials.py:1576-1601)
        uses_direction = bool(getattr(sensor.material, "uses_direction", True))
        planes = self.collect_grids(grids, w_edges, device=device, reduce=reduce, moments=uses_direction)
        flux, moment = planes["flux"], planes.get("moment_real")
        axes_out = tuple(compiled.shape) + (axis_wavelength, sensor.axis_pixel.x, sensor.axis_pixel.y)
        from .vectors import SpectralPositionalVectorArray

        inputs = SpectralPositionalVectorArray(
            wavelength=na.ScalarArray(w_edges, axis_wavelength),
            position=na.Cartesian2dVectorArray(
                x=na.ScalarArray(ex, (sensor.axis_pixel.x,)), y=na.ScalarArray(ey, (sensor.axis_pixel.y,))
            ),
        )
        collected = na.FunctionArray(inputs=inputs, outputs=na.ScalarArray(flux, axes_out))

        def direction():  # sensors/_sensors.py:163-169; only materials that depend on the angle of incidence ask
            with np.errstate(invalid="ignore", divide="ignore"):
                return na.ScalarArray(np.where(flux > 0, moment / flux, 1) + 0j, axes_out)

        return sensor.expose(collected, direction, axis_wavelength=axis_wavelength, noise=noise, seed=seed)
