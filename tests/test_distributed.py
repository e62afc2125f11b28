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


# ---------------------------------------------------------------------------
# ImagePipeline: per-configuration reduce + read-back into a shared host buffer
# ---------------------------------------------------------------------------
def _pipeline_worker(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from optika_b200 import _engine

    ew, ex, ey = np.array([0.0, 1.0, 2.0]), np.linspace(-1, 1, 6), np.linspace(-1, 1, 4)  # 2 x 5 x 3 = 30 bins
    image = _engine.DeviceImage.zeros(ew, ex, ey, "cpu", leading=(3,), moments=True, counts=True, fused=True, pad_to=world)
    assert image.buffer_f64.shape == (3, 60) and image.buffer_i64.shape == (3, 30)
    pipeline = distributed.ImagePipeline(image, "cpu")
    results = []
    for exposure in range(2):  # the buffers are reused
        image.zero_()
        for c in range(3):
            image.flux[c] += (rank + 1) * (c + 1) * torch.arange(30.0, dtype=torch.float64).reshape(2, 5, 3)
            image.moment_real[c] += 0.5 * (rank + 1)
            image.counts[c] += (rank + 1) * (exposure + 1)
            pipeline.submit(c)
        planes = pipeline.finish()
        results.append({k: np.array(v) for k, v in planes.items()})
        dist.barrier()
    if rank == 0:
        queue.put(results)
    dist.barrier()
    pipeline.close()
    dist.destroy_process_group()


def test_image_pipeline_reduces_into_a_shared_host_buffer_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    base = np.arange(30.0).reshape(2, 5, 3)
    for exposure, planes in enumerate(results):
        assert planes["flux"].shape == (3, 2, 5, 3)
        for c in range(3):
            assert np.array_equal(planes["flux"][c], 3 * (c + 1) * base)  # ranks contribute 1x and 2x
            assert np.array_equal(planes["moment_real"][c], np.full((2, 5, 3), 1.5))
            assert np.array_equal(planes["counts"][c], np.full((2, 5, 3), 3 * (exposure + 1)))


def test_image_pipeline_single_process_is_a_plain_read_back():
    from optika_b200 import _engine

    image = _engine.DeviceImage.zeros(np.array([0.0, 1.0]), np.linspace(0, 1, 4), np.linspace(0, 1, 3), "cpu",
                                      moments=True, counts=False, fused=True)
    image.flux += 2.0
    image.moment_real += 3.0
    pipeline = distributed.ImagePipeline(image, "cpu")
    pipeline.submit(0)
    planes = pipeline.finish()
    assert planes["flux"].shape == (1, 3, 2) and np.all(planes["flux"] == 2.0) and np.all(planes["moment_real"] == 3.0)
    assert "counts" not in planes
    pipeline.close()
    # one plane: what `SequentialSystem.image` asks for when the sensor material ignores the angle of incidence
    lean = _engine.DeviceImage.zeros(np.array([0.0, 1.0]), np.linspace(0, 1, 4), np.linspace(0, 1, 3), "cpu",
                                     leading=(2,), moments=False, counts=False, fused=True, pad_to=4)
    assert lean.moment_real is None and lean.buffer_f64.shape == (2, 8) and lean.buffer_i64 is None
    lean.flux[1] += 5.0
    pipeline = distributed.ImagePipeline(lean, "cpu")
    for c in range(2):
        pipeline.submit(c)
    planes = pipeline.finish()
    assert set(planes) == {"flux"} and planes["flux"].shape == (2, 1, 3, 2)
    assert np.all(planes["flux"][0] == 0.0) and np.all(planes["flux"][1] == 5.0)
    pipeline.close()


def test_fused_image_planes_are_addressed_through_their_strides():
    from optika_b200 import _engine

    image = _engine.DeviceImage.zeros(np.array([0.0, 1.0]), np.linspace(0, 1, 5), np.linspace(0, 1, 3), "cpu",
                                      leading=(2, 3), moments=True, counts=True, fused=True, pad_to=8)
    n = 4 * 2
    assert image.buffer_f64.shape == (6, 16) and image.buffer_i64.shape == (6, 8)
    for c in range(6):
        im = image.struct(c)
        assert im.flux == image.buffer_f64[c].data_ptr()
        assert im.moment_real == image.buffer_f64[c].data_ptr() + 8 * n
        assert im.counts == image.buffer_i64[c].data_ptr()
    plain = _engine.DeviceImage.zeros(np.array([0.0, 1.0]), np.linspace(0, 1, 5), np.linspace(0, 1, 3), "cpu", leading=(2, 3))
    assert plain.struct(4).flux == plain.flux.data_ptr() + 4 * n * 8


def test_best_shard_axis_balances_the_slabs():
    assert distributed.best_shard_axis((1, 354, 354, 100, 100), 8) == 1  # 354 / 8: 1.7 % imbalance, 100 / 8: 4 %
    assert distributed.best_shard_axis((1, 100, 100, 112, 112), 8) == 3  # 112 = 8 x 14 exactly: pupil x
    assert distributed.best_shard_axis((1, 4096, 4096, 10, 8), 8) == 1   # 10 pupil cells do not split into 8
    assert distributed.best_shard_axis((1, 4, 4, 100, 100), 1) == 3
    assert distributed.best_shard_axis((1, 3, 3, 2, 64), 8) == 4         # the only axis with enough cells
