"""
The sequential loop over surfaces.

Mirrors ``optika.propagators.propagate_rays`` / ``accumulate_rays``
(``optika/propagators.py:19-73``) -- the choke point of the reference's hot path.
Instead of a Python ``for`` over surfaces with ~150 NumPy passes each, the whole
list is lowered to a device table and every ray walks it inside one kernel
launch (``optk_trace``).  Inputs and outputs keep the reference's named axes;
host arrays come back as host arrays.
"""

from __future__ import annotations
from typing import Sequence
from . import _engine
from . import _lib as L

__all__ = ["propagate_rays", "accumulate_rays"]


def _as_list(propagators) -> list:
    if hasattr(propagators, "propagate_rays") and not isinstance(propagators, (list, tuple)):
        return [propagators]
    return list(propagators)


def _chain(surfaces: list, rays, accumulate: bool, axis: str | None, device):
    """Trace through any number of surfaces, chaining launches of <= OPTK_MAX_SURFACES."""
    if not accumulate:
        for k in range(0, len(surfaces), L.MAX_SURFACES):
            system = _engine.CompiledSystem(surfaces[k : k + L.MAX_SURFACES])
            rays = _engine.trace(system, rays, device=device)
        return rays
    if len(surfaces) <= L.MAX_SURFACES:
        system = _engine.CompiledSystem(surfaces)
        return _engine.trace(system, rays, accumulate=True, axis=axis, device=device)
    # longer lists: every link is lowered over the configuration shape of the WHOLE list (so all links
    # produce the same axes), traced with accumulate, and fed the last state of the link before it
    from . import _lowering

    shape_ = _lowering.config_shape(surfaces)
    parts = []
    for k in range(0, len(surfaces), L.MAX_SURFACES):
        link = surfaces[k : k + L.MAX_SURFACES]
        system = _engine.CompiledSystem(link, lowered=_lowering.lower_system(link, shape_=dict(shape_)))
        out = _engine.trace(system, rays, accumulate=True, axis=axis, device=device)
        parts.append(out)
        rays = out.last_state()
    return _engine.DeviceRays.concatenate(parts, axis if axis is not None else "surface")


def propagate_rays(propagators, rays, device=None):
    """
    Propagate `rays` through every surface in `propagators`
    (``optika/propagators.py:19-41``).  A :class:`~optika_b200.rays.RayVectorArray`
    comes back as one (host); :class:`~optika_b200._engine.DeviceRays` stay in HBM.
    """
    surfaces = _as_list(propagators)
    on_device = isinstance(rays, _engine.DeviceRays)
    if not surfaces:
        return rays
    result = _chain(surfaces, rays, False, None, device)
    return result if on_device else result.to_host()


def accumulate_rays(propagators, rays, axis: str, device=None):
    """
    Like :func:`propagate_rays` but keeps the rays at every surface on the new
    named axis `axis` (``optika/propagators.py:44-73``).
    """
    surfaces = _as_list(propagators)
    on_device = isinstance(rays, _engine.DeviceRays)
    result = _chain(surfaces, rays, True, axis, device)
    return result if on_device else result.to_host()
