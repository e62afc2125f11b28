"""
The N > 1 path on CPU: world_size-2 ``gloo`` process group.  Rays are sharded by
pupil slab with no data-path collective; the detector image planes are summed
with one all-reduce (integer counts exactly, fp64 sums to rounding).
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import configs
from optika_b200 import distributed
from oracle import raytrace as ora, binning as orb


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_image(system, grid):
    """Oracle trace + histogram of one slab of the grid (CPU stand-in for the device path)."""
    _, rays = system._calc_rayfunction_input(grid).inputs, system._calc_rayfunction_input(grid).outputs
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    out = ora.propagate_rays(system.surfaces_all, r0)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    ew = np.array([1e-5, 1e-4])
    counts = orb.counts(local, ew, ex, ey)
    flux, _, _ = orb.collect(local, ew, ex, ey)
    return counts, flux


def _worker(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    system = configs.spherical_grating(num_field=3, num_pupil=12, num_wavelength=3, num_pixel=64)
    mine = distributed.shard_grid(system.grid_input, "pupil_x", rank, world)
    counts, flux = _oracle_image(system, mine)
    planes = [torch.from_numpy(counts.copy()), torch.from_numpy(flux.copy())]
    distributed.reduce_image(planes)
    if rank == 0:
        queue.put((planes[0].numpy(), planes[1].numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_image_equals_whole_image_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    counts, flux = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    system = configs.spherical_grating(num_field=3, num_pupil=12, num_wavelength=3, num_pixel=64)
    want_counts, want_flux = _oracle_image(system, system.grid_input)
    assert np.array_equal(counts, want_counts)  # integer counts reduce exactly
    assert np.allclose(flux, want_flux, rtol=1e-12)
    assert counts.sum() > 0


def test_reduce_image_is_identity_without_process_group():
    t = torch.arange(4.0)
    distributed.reduce_image([t])
    assert torch.equal(t, torch.arange(4.0))


def _grid_vertices():
    return [
        np.linspace(30e-6, 50e-6, 3), np.linspace(-5e-4, 5e-4, 4), np.linspace(-5e-4, 5e-4, 4),
        np.linspace(-45, 45, 11), np.linspace(-45, 45, 9),
    ]


def _oracle_grid_image(system, grid, seed):
    """Oracle image of the sub-box of an on-device ray grid (CPU stand-in for optk_trace_grid)."""
    from oracle import grid as og

    rays0 = og.input_rays(list(grid.vertices), begin=grid.begin, count=grid.count, seed=seed)
    out = ora.propagate_rays(system.surfaces_all, rays0)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    ew = np.array([1e-5, 1e-4])
    return orb.counts(local, ew, ex, ey), orb.collect(local, ew, ex, ey)[0]


def _grid_worker(rank, world, port, queue):
    from optika_b200 import _grid

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert distributed.rank_world() == (rank, world)
    system = configs.spherical_grating(num_field=3, num_pupil=12, num_wavelength=3, num_pixel=64)
    mine = _grid.RayGrid(_grid_vertices(), seed=21).shard(*distributed.rank_world())  # what image() does per rank
    counts, flux = _oracle_grid_image(system, mine, seed=21)
    planes = [torch.from_numpy(counts.copy()), torch.from_numpy(flux.copy())]
    distributed.reduce_image(planes)
    if rank == 0:
        queue.put((planes[0].numpy(), planes[1].numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ray_grid_equals_whole_grid_gloo():
    """
    The random stream of the on-device grid is keyed by the cell index in the WHOLE grid, so
    the slabs traced by the ranks sum to exactly the single-process image.
    """
    from optika_b200 import _grid

    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    counts, flux = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    system = configs.spherical_grating(num_field=3, num_pupil=12, num_wavelength=3, num_pixel=64)
    want_counts, want_flux = _oracle_grid_image(system, _grid.RayGrid(_grid_vertices(), seed=21), seed=21)
    assert np.array_equal(counts, want_counts) and counts.sum() > 0
    assert np.allclose(flux, want_flux, rtol=1e-12)
    assert distributed.rank_world() == (0, 1)
