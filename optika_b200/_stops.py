"""
The stop solver: which rays connect the pupil stop and the field stop.

Mirrors ``SequentialSystem._calc_rayfunction_stops_only`` / ``_calc_rayfunction_stops``
/ ``_denormalize_grid`` (``optika/systems/_sequential.py:396-678, 748-789``).  This is a
host-side *caller* of the hot path: a 2-D Newton
iteration with a finite-difference Jacobian whose residual function traces a
small batch of rays through the sub-system between the two stop surfaces.  With
the device backend the whole iteration runs in one kernel launch per
configuration (``optk_solve_stops``, one thread per unknown ray); the host
iteration below (every trace on the device through ``propagators.propagate_rays``,
Newton bookkeeping in NumPy, exactly where the reference keeps it) remains for
injected backends and for the cases the kernel does not cover.

The two primitives the solver needs -- ``propagate(surfaces, rays)`` and
``sag(surface, x, y)`` -- come from a small backend object (default: the device
engine), so the tests can drive the very same solver with the NumPy oracle and
compare the converged rays.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na
from . import units as u
from . import _util
from .rays import RayVectorArray
from .vectors import ObjectVectorArray

__all__ = ["rayfunction_stops", "denormalize_grid", "AXIS_FIELD_STOP", "AXIS_PUPIL_STOP"]

AXIS_PUPIL_STOP = "_stop_pupil"  # _sequential.py:680-681
AXIS_FIELD_STOP = "_stop_field"


class DeviceBackend:
    """The device engine: every trace and sag evaluation runs in the CUDA kernels."""

    @staticmethod
    def propagate(surfaces, rays):
        from . import propagators

        return propagators.propagate_rays(surfaces, rays)

    @staticmethod
    def sag(surface, x, y):
        return surface.sag(na.Cartesian3dVectorArray(x, y, 0.0 * (x + y)))

    @staticmethod
    def solve(surfaces, rays, variable, x0, y0, target_xy, target, step, max_abs_error):
        """The whole Newton iteration on the device (``optk_solve_stops``); ``None`` = not covered."""
        from . import _engine

        return _engine.solve_stops(surfaces, rays, variable, x0, y0, target_xy, target, step, max_abs_error)


def _is_angular(aperture) -> bool:
    return bool(getattr(aperture, "angular", False))


def _moveaxis(v: na.Cartesian3dVectorArray, source: str, destination: str) -> na.Cartesian2dVectorArray:
    def rename(c):
        c = na.as_named_array(c)
        return na.ScalarArray(c.ndarray, tuple(destination if ax == source else ax for ax in c.axes))

    return na.Cartesian2dVectorArray(rename(v.x), rename(v.y))


def _sag_z(surface, x, y, backend) -> na.ScalarArray:
    """``surface.sag(position)`` at the points ``(x, y)``."""
    name = type(surface.sag).__name__
    if name == "NoSag" and surface.sag.transformation is None:
        return 0.0 * (x + y)  # optika/sags/_flat.py:26-41 without a transformation
    return backend.sag(surface, x, y)


def _anchor_surface(subsystem):
    # _sequential.py:343-361
    for surface in subsystem[1:]:
        material = surface.material
        if material is not None and material.is_mirror:
            return surface
        if surface.sag is not None and type(surface.sag).__name__ != "NoSag":
            return surface
        if surface.rulings is not None:
            return surface
    return subsystem[-1]


def _stop_indices(system):
    surfaces = system.surfaces_all
    if not any(s.is_field_stop for s in surfaces):
        # _sequential.py:109-110: without an explicit field stop the first surface is one
        surfaces = [dataclasses.replace(surfaces[0], is_field_stop=True)] + surfaces[1:]
    pupil = [i for i, s in enumerate(surfaces) if s.is_pupil_stop]
    field = [i for i, s in enumerate(surfaces) if s.is_field_stop]
    if not pupil:
        raise ValueError(
            "Pupil stop is not defined for this system."
            "Set `is_pupil_stop=True` for at least one surface in this system."
        )
    if not field:
        raise ValueError("field stop not defined")
    return surfaces, pupil, field


def _newton(function, x, y, dx, max_abs_error, max_iterations=100):
    """
    ``na.optimize.root_newton`` with the finite-difference ``na.jacobian`` the
    reference passes (``_sequential.py:586-606``; third-party named_arrays ~= 2.1):
    restated as the textbook 2-D Newton iteration, forward differences of step `dx`,
    iterated until every component of the residual is below `max_abs_error`.
    """
    for _ in range(max_iterations):
        f = function(x, y)
        if np.all(np.abs(f.x.ndarray) <= max_abs_error) and np.all(np.abs(f.y.ndarray) <= max_abs_error):
            return x, y
        fx = function(x + dx, y)
        fy = function(x, y + dx)
        j11, j21 = (fx.x - f.x) / dx, (fx.y - f.y) / dx
        j12, j22 = (fy.x - f.x) / dx, (fy.y - f.y) / dx
        det = j11 * j22 - j12 * j21
        x = x - (j22 * f.x - j12 * f.y) / det
        y = y - (-j21 * f.x + j11 * f.y) / det
    raise ValueError("Max iterations exceeded")


def _stops_only(system, wavelength, samples_pupil_stop, samples_field_stop, backend):
    """``_calc_rayfunction_stops_only``, ``_sequential.py:396-623``."""
    surfaces, indices_pupil, indices_field = _stop_indices(system)
    inputs = ObjectVectorArray(wavelength=wavelength)
    rays = RayVectorArray(wavelength=u.length(wavelength))
    while indices_pupil and indices_field:
        index_first = min(indices_pupil[0], indices_field[0])
        index_last = max(indices_pupil[0], indices_field[0])
        subsystem = surfaces[index_first : index_last + 1]
        surface_first, surface_last = subsystem[0], subsystem[-1]
        first_is_pupil = surface_first.is_pupil_stop
        axis_first = AXIS_PUPIL_STOP if first_is_pupil else AXIS_FIELD_STOP
        axis_last = AXIS_FIELD_STOP if first_is_pupil else AXIS_PUPIL_STOP
        grid_first = _moveaxis(
            surface_first.aperture.wire(samples_pupil_stop if first_is_pupil else samples_field_stop),
            "wire", axis_first,
        )
        grid_last = _moveaxis(
            surface_last.aperture.wire(samples_field_stop if first_is_pupil else samples_pupil_stop),
            "wire", axis_last,
        )
        if first_is_pupil:
            indices_pupil.pop(0)
            inputs.pupil, inputs.field = grid_first, grid_last
        else:
            indices_field.pop(0)
            inputs.field, inputs.pupil = grid_first, grid_last

        first_angular = _is_angular(surface_first.aperture)
        last_angular = _is_angular(surface_last.aperture)
        if not first_angular:
            position = na.Cartesian3dVectorArray(
                grid_first.x, grid_first.y, _sag_z(surface_first, grid_first.x, grid_first.y, backend)
            )
            direction = na.Cartesian3dVectorArray(0.0, 0.0, 1.0)
            variable = "direction"

            def zfunc(x, y):
                return np.sqrt(1 - (np.square(x) + np.square(y)))
        else:
            direction = na.Cartesian3dVectorArray(
                grid_first.x, grid_first.y, np.sqrt(1 - (np.square(grid_first.x) + np.square(grid_first.y)))
            )
            position = na.Cartesian3dVectorArray(0.0, 0.0, 0.0)
            variable = "position"

            def zfunc(x, y, _s=surface_first):
                return _sag_z(_s, x, y, backend)

        # seed (_sequential.py:497-549)
        anchor = _anchor_surface(subsystem)
        if anchor is surface_last and not last_angular:
            aim = na.Cartesian3dVectorArray(
                grid_last.x, grid_last.y, _sag_z(surface_last, grid_last.x, grid_last.y, backend)
            )
            if surface_last.transformation is not None:
                aim = surface_last.transformation(aim)
        else:
            aim = na.Cartesian3dVectorArray(0.0, 0.0, 0.0)
            if anchor.transformation is not None:
                aim = anchor.transformation(aim)
        if surface_first.transformation is not None:
            aim = surface_first.transformation.inverse(aim)
        if variable == "direction":
            d = aim - position
            d = d / d.length
            flip = np.sign(d.z)
            where = d.z != 0
            direction = na.Cartesian3dVectorArray(
                x=np.where(where, flip * d.x, 0.0),
                y=np.where(where, flip * d.y, 0.0),
                z=np.where(where, flip * d.z, 1.0),
            )
        else:
            t = aim.z / direction.z
            sx, sy = aim.x - direction.x * t, aim.y - direction.y * t
            position = na.Cartesian3dVectorArray(sx, sy, _sag_z(surface_first, sx, sy, backend))

        rays = dataclasses.replace(rays, position=position, direction=direction)
        # from here on the rays live in GLOBAL coordinates and the Newton variables are the
        # global x, y components of the free vector (_sequential.py:551-553, 570)
        if surface_first.transformation is not None:
            rays = surface_first.transformation(rays)

        scale = max(float(grid_last.x.ptp().ndarray), float(grid_last.y.ptp().ndarray))
        max_abs_error = 1e-9 * max(scale, 1.0)  # _sequential.py:561-568
        dx = 1e-6 if variable == "direction" else 1e-6 * max(scale, 1.0)  # :586-591
        target = "direction" if last_angular else "position"
        transformation_last = surface_last.transformation

        def function(x, y, _rays=rays):
            # _ray_error, _sequential.py:363-394
            vec = na.Cartesian3dVectorArray(x, y, zfunc(x, y))
            trial = backend.propagate(subsystem[1:], dataclasses.replace(_rays, **{variable: vec}))
            if transformation_last is not None:
                trial = transformation_last.inverse(trial)
            got = getattr(trial, target)
            return na.Cartesian2dVectorArray(got.x - grid_last.x, got.y - grid_last.y)

        variables = getattr(rays, variable)
        shape_ = na.shape_broadcasted(variables.x, variables.y, grid_first, grid_last, rays.wavelength)
        x0 = na.broadcast_to(na.as_named_array(variables.x), shape_).copy()
        y0 = na.broadcast_to(na.as_named_array(variables.y), shape_).copy()
        try:
            # the device backend runs the whole iteration in one launch per configuration
            # (SURVEY.md section 8f-1); other backends, and problems the kernel does not cover,
            # iterate here with one batch of traces per residual evaluation
            solve = getattr(backend, "solve", None)
            solved = None
            if solve is not None:
                solved = solve(subsystem, rays, variable, x0, y0, grid_last, target, dx, max_abs_error)
            if solved is None:
                rx, ry = _newton(function, x0, y0, dx, max_abs_error)
                solved = (rx, ry, zfunc(rx, ry))
        except ValueError as e:
            raise ValueError(
                f"Could not solve for the rays connecting the stop surfaces "
                f"{surface_first.name!r} and {surface_last.name!r}."
            ) from e
        rays = dataclasses.replace(rays, **{variable: na.Cartesian3dVectorArray(*solved)})
    return inputs, rays, surfaces


def rayfunction_stops(system, wavelength, samples_pupil_stop=101, samples_field_stop=101, backend=None):
    """
    ``_calc_rayfunction_stops`` (``_sequential.py:625-678``): the stop rays propagated
    BACKWARDS from the first stop to the object surface (``surfaces[index_stop::-1]``).
    Returns ``(inputs, rays)`` with the rays in the object's global coordinates.
    """
    backend = backend or DeviceBackend
    inputs, rays, surfaces = _stops_only(system, wavelength, samples_pupil_stop, samples_field_stop, backend)
    index_pupil = [i for i, s in enumerate(surfaces) if s.is_pupil_stop][-1]
    index_field = [i for i, s in enumerate(surfaces) if s.is_field_stop][-1]
    index_stop = min(index_pupil, index_field)
    subsystem = surfaces[index_stop::-1]
    rays = backend.propagate(subsystem, rays)
    obj = subsystem[-1]
    local = obj.transformation.inverse(rays) if obj.transformation is not None else rays
    # where = direction @ sag.normal(position) > 0: flip the direction (:656-657); the object is flat
    normal_z = -1.0
    shape_ = na.shape_broadcasted(rays.position, rays.direction)
    flip = np.where(na.broadcast_to(na.as_named_array(local.direction.z * normal_z), shape_).ndarray > 0, -1.0, 1.0)
    flip = na.ScalarArray(flip, tuple(shape_))
    rays = dataclasses.replace(
        rays,
        position=rays.position.broadcast_to(shape_),
        direction=na.Cartesian3dVectorArray(
            *[na.broadcast_to(na.as_named_array(c), shape_) * flip for c in rays.direction.components]
        ),
    )
    if system.transformation is not None:
        rays = system.transformation(rays)
    return inputs, rays


def stop_extents(system, wavelength, backend=None) -> tuple:
    """
    ``(field_min, field_ptp, pupil_min, pupil_ptp)`` of the rays through the edges of both stops
    (21 x 21 samples), the quantities ``_denormalize_grid`` scales with (``_sequential.py:761-787``).
    """
    _, rays = rayfunction_stops(system, wavelength, samples_pupil_stop=21, samples_field_stop=21, backend=backend)
    axes = (AXIS_FIELD_STOP, AXIS_PUPIL_STOP)
    if system.object_is_at_infinity:
        field = _util.angles(rays.direction)
        pupil = na.Cartesian2dVectorArray(rays.position.x, rays.position.y)
    else:
        field = na.Cartesian2dVectorArray(rays.position.x, rays.position.y)
        pupil = _util.angles(rays.direction)
    return field.min(axis=axes), field.ptp(axis=axes), pupil.min(axis=axes), pupil.ptp(axis=axes)


def denormalize_grid(system, grid: ObjectVectorArray, normalized_field=True, normalized_pupil=True, backend=None,
                     extents=None):
    """``_denormalize_grid`` (``_sequential.py:748-789``): normalised [-1, 1] -> physical coordinates."""
    if (not normalized_field) and (not normalized_pupil):
        return grid
    if extents is None:
        extents = stop_extents(system, grid.wavelength, backend)
    field_lo, field_ptp, pupil_lo, pupil_ptp = extents
    result = grid.copy_shallow()
    if normalized_field:
        result.field = field_ptp * (result.field + 1) / 2 + field_lo
    if normalized_pupil:
        result.pupil = pupil_ptp * (result.pupil + 1) / 2 + pupil_lo
    return result
