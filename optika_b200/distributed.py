"""
Multi-GPU: one process per GPU, rays sharded by contiguous field / pupil slab.

The reference has no multi-process mode; rays are
independent, so the trace needs no exchange at all.  The only collective on the
path is the sum of the detector image planes after binning
(``optika/sensors/_sensors.py:155-161`` summed over ranks), done with
``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).
Integer hit counts reduce exactly; weighted fp64 sums to rounding.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na
from .vectors import ObjectVectorArray

__all__ = ["slab", "shard_axis", "shard_grid", "reduce_image", "image_rays_sharded", "rank_world"]


def rank_world() -> tuple[int, int]:
    """(rank, world size) of the default process group; (0, 1) outside ``torch.distributed``."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def slab(n: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slab of ``range(n)`` owned by `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} out of range for world size {world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def shard_axis(a, axis: str, rank: int, world: int):
    """The slab of named array (or vector) `a` along `axis`; objects without `axis` pass through."""
    if isinstance(a, (na.Cartesian2dVectorArray, na.Cartesian3dVectorArray)):
        return a._map(lambda c: shard_axis(c, axis, rank, world))
    if isinstance(a, na.ScalarArray) and axis in a.axes:
        return a[{axis: slab(a.shape[axis], rank, world)}]
    return a


def shard_grid(grid: ObjectVectorArray, axis: str, rank: int, world: int) -> ObjectVectorArray:
    """This rank's slab of an input grid along the named `axis` (e.g. ``"pupil_x"``)."""
    return ObjectVectorArray(
        wavelength=shard_axis(grid.wavelength, axis, rank, world),
        field=shard_axis(grid.field, axis, rank, world),
        pupil=shard_axis(grid.pupil, axis, rank, world),
    )


def reduce_image(planes, group=None):
    """
    Sum detector planes over all ranks, in place (``all_reduce``).  `planes` is a
    :class:`~optika_b200._engine.DeviceImage` or an iterable of tensors.
    """
    import torch.distributed as dist

    if hasattr(planes, "flux"):
        tensors = [t for t in (planes.flux, planes.moment_real, planes.moment_imag, planes.counts) if t is not None]
    else:
        tensors = list(planes)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return planes


def image_rays_sharded(system, wavelength_edges, axis: str = "pupil_x", grid=None, device=None, **kwargs):
    """
    Detector image of the whole input grid with the rays sharded along `axis`
    over the ranks of the default process group: every rank traces and bins its
    slab (fused kernel), then the planes are summed with one all-reduce.
    """
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    grid = system.grid_input if grid is None else grid
    mine = shard_grid(grid, axis, rank, world)
    image = system.image_rays(
        wavelength_edges, wavelength=mine.wavelength, field=mine.field, pupil=mine.pupil, device=device, **kwargs
    )
    return reduce_image(image)
