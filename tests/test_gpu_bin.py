"""
Parity of detector binning (kernel 2 and the fused trace+bin path) against the
oracle's numpy.histogramdd restatement of ``optika/sensors/_sensors.py:92-171``.
Pixel COUNTS must be bit-exact except for rays within tolerance of a bin edge
(enumerated); weighted sums within 1e-9 (atomic summation order differs).
"""

import ctypes as C

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import _engine, _lib
from oracle import raytrace as ora, binning as orb

import configs

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(2)


def sensor(nx=64, ny=48, width=15 * u.um):
    return optika.sensors.ImagingSensor(
        width_pixel=width,
        axis_pixel=na.Cartesian2dVectorArray("detector_x", "detector_y"),
        num_pixel=na.Cartesian2dVectorArray(nx, ny),
    )


def random_local_rays(n, s, spread=1.2):
    ex, ey = s.pixel_edges()
    ax = "ray"
    x = rng.uniform(ex[0] * spread, ex[-1] * spread, n)
    y = rng.uniform(ey[0] * spread, ey[-1] * spread, n)
    w = rng.uniform(90 * u.AA, 310 * u.AA, n)
    rays = optika.rays.RayVectorArray(
        wavelength=na.ScalarArray(w, ax),
        position=na.Cartesian3dVectorArray(na.ScalarArray(x, ax), na.ScalarArray(y, ax), 0.0),
        direction=na.Cartesian3dVectorArray(0.0, 0.0, na.ScalarArray(rng.uniform(0.8, 1.0, n), ax)),
        intensity=na.ScalarArray(rng.uniform(0.0, 2.0, n), ax),
        unvignetted=na.ScalarArray(rng.uniform(size=n) > 0.2, ax),
    )
    return rays


def test_collect_matches_histogramdd(cuda_device):
    s = sensor()
    rays = random_local_rays(200_000, s)
    edges = na.ScalarArray(np.linspace(100, 300, 6) * u.AA, "wavelength")
    image, direction = s.collect(rays, wavelength=edges, axis="ray")
    r0, _ = configs.flatten_rays(rays)
    ex, ey = s.pixel_edges()
    want, want_direction, _ = orb.collect(r0, edges.ndarray, ex, ey)
    assert image.outputs.shape == {"wavelength": 5, "detector_x": 64, "detector_y": 48}
    assert np.allclose(image.outputs.ndarray, want, rtol=1e-9, atol=1e-12)
    assert np.allclose(direction.ndarray, want_direction, rtol=1e-9)
    assert np.array_equal(image.inputs.position.x.ndarray, ex)


def test_counts_bit_exact_with_edge_samples(cuda_device):
    """Samples placed EXACTLY on bin edges, on the last edge, outside, and NaN."""
    s = sensor(nx=16, ny=8)
    ex, ey = s.pixel_edges()
    ew = np.array([1.0e-5, 2.0e-5, 4.0e-5])
    xs = np.concatenate([ex, ex[:-1] + np.diff(ex) / 2, [ex[0] - 1e-9, ex[-1] + 1e-9, np.nan, np.inf]])
    ys = np.concatenate([ey, ey[:-1] + np.diff(ey) / 2, [ey[0] - 1e-9, ey[-1] + 1e-9, np.nan]])
    ws = np.array([1.0e-5, 1.5e-5, 2.0e-5, 4.0e-5, 4.0000001e-5, 0.9e-5, np.nan])
    X, Y, W = np.meshgrid(xs, ys, ws, indexing="ij")
    n = X.size
    image = _engine.DeviceImage.zeros(ew, ex, ey, cuda_device, moments=True, counts=True)
    import torch

    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.reshape(-1))).to(cuda_device)  # noqa: E731
    x, y, w = dev(X), dev(Y), dev(W)
    inten = torch.ones(n, dtype=torch.float64, device=cuda_device)
    im = image.struct(0)
    _lib.check(
        _lib.lib().optk_bin(
            n, w.data_ptr(), x.data_ptr(), y.data_ptr(), None, inten.data_ptr(), None, C.byref(im),
            torch.cuda.current_stream().cuda_stream,
        )
    )
    rays = dict(
        wavelength=W.reshape(-1), px=X.reshape(-1), py=Y.reshape(-1), dz=np.ones(n), intensity=np.ones(n),
        unvignetted=np.ones(n, dtype=bool),
    )
    want = orb.counts(rays, ew, ex, ey)
    got = image.counts.cpu().numpy()
    assert np.array_equal(got, want)
    assert want.sum() > 0
    assert np.array_equal(image.flux.cpu().numpy(), want.astype(float))


def test_fused_trace_bin_equals_trace_then_collect(cuda_device):
    """image_rays (no ray ever written to HBM) == rayfunction + collect == oracle."""
    system = configs.newtonian(num_field=6, num_pupil=24, num_pixel=128)
    edges = na.ScalarArray(np.array([499.0, 501.0]) * u.nm, "wavelength")
    fused = system.image_rays(edges, **configs.PHYSICAL)
    rays = system.rayfunction(on_device=True, **configs.PHYSICAL).outputs
    image, _ = system.sensor.collect(rays, wavelength=edges)
    assert np.allclose(fused.flux.cpu().numpy().reshape(image.outputs.ndarray.shape), image.outputs.ndarray, rtol=1e-12)
    # oracle
    _, rays_in = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays_in)
    out = ora.propagate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    want = orb.counts(local, edges.ndarray, ex, ey)
    got = fused.counts.cpu().numpy().reshape(want.shape)
    mism = got != want
    if mism.any():
        # enumerate rays within 1e-9 relative of a bin edge; they alone may move between bins
        near = (orb.bin_margin(local["px"], ex) < 1e-9 * abs(ex[0])) | (orb.bin_margin(local["py"], ey) < 1e-9 * abs(ey[0]))
        assert mism.sum() <= 2 * near.sum()
    assert got.sum() == want.sum()
    flux, _, _ = orb.collect(local, edges.ndarray, ex, ey)
    if not mism.any():
        assert np.allclose(fused.flux.cpu().numpy().reshape(flux.shape), flux, rtol=1e-9)


def test_fused_image_with_explicit_frame_equals_local_trace(cuda_device):
    """optk_trace(image, image_frame = sensor.transformation) on the GLOBAL trace == binning the local trace."""
    system = configs.newtonian(num_field=4, num_pupil=20, num_pixel=64)
    edges = na.ScalarArray(np.array([499.0, 501.0]) * u.nm, "wavelength")
    local = system.image_rays(edges, **configs.PHYSICAL)
    _, rays = system._input(None, None, None, None, False, False)
    ex, ey = system.sensor.pixel_edges()
    image = _engine.DeviceImage.zeros(edges.ndarray, ex, ey, cuda_device, moments=True, counts=True)
    _engine.trace(
        system._compiled, rays, image=image, image_frame=system.sensor.transformation, write_rays=False,
        ray_axes_order=system._ray_axes_order,
    )
    a, b = local.counts.cpu().numpy().reshape(-1), image.counts.cpu().numpy().reshape(-1)
    assert a.sum() == b.sum() > 0
    assert (a != b).sum() <= 4  # a ray within rounding of a pixel edge may move
    assert np.allclose(local.flux.sum().item(), image.flux.sum().item(), rtol=1e-12)


def test_fused_image_with_configuration_axis(cuda_device):
    system = configs.misaligned_telescope(num_field=3, num_pupil=12, num_pixel=64, num_tilt=3)
    edges = na.ScalarArray(np.array([499.0, 501.0]) * u.nm, "wavelength")
    image = system.image_rays(edges, **configs.PHYSICAL)
    counts = image.counts.cpu().numpy()
    assert counts.shape == (3, 1, 64, 64)
    _, rays_in = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays_in)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    ex, ey = system.sensor.pixel_edges()
    for i in range(3):
        surfaces = [ora.select_config(s, {"misalign": i}) for s in system.surfaces_all]
        out = ora.propagate_rays(surfaces, r0, extended=True)
        local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
        want = orb.counts(local, edges.ndarray, ex, ey)
        assert counts[i].sum() == want.sum()
        assert (counts[i] != want).sum() <= 4
    assert not np.array_equal(counts[0], counts[2])


def test_size_independent_properties_at_scale(cuda_device):
    """1e7 rays: every unvignetted in-range ray is counted exactly once; flux sums match."""
    import torch

    n = 10_000_000
    s = sensor(nx=512, ny=512)
    ex, ey = s.pixel_edges()
    g = torch.Generator(device=cuda_device).manual_seed(0)
    x = (torch.rand(n, generator=g, device=cuda_device, dtype=torch.float64) - 0.5) * 2.2 * ex[-1]
    y = (torch.rand(n, generator=g, device=cuda_device, dtype=torch.float64) - 0.5) * 2.2 * ey[-1]
    w = torch.full((n,), 2e-5, dtype=torch.float64, device=cuda_device)
    inten = torch.rand(n, generator=g, device=cuda_device, dtype=torch.float64)
    mask = (torch.rand(n, generator=g, device=cuda_device) > 0.3).to(torch.uint8)
    image = _engine.DeviceImage.zeros(np.array([1e-5, 3e-5]), ex, ey, cuda_device, moments=False, counts=True)
    im = image.struct(0)
    _lib.check(
        _lib.lib().optk_bin(
            n, w.data_ptr(), x.data_ptr(), y.data_ptr(), None, inten.data_ptr(), mask.data_ptr(), C.byref(im),
            torch.cuda.current_stream().cuda_stream,
        )
    )
    inside = (x >= ex[0]) & (x <= ex[-1]) & (y >= ey[0]) & (y <= ey[-1]) & (mask != 0)
    assert int(image.counts.sum().item()) == int(inside.sum().item())
    assert np.isclose(float(image.flux.sum().item()), float(inten[inside].sum().item()), rtol=1e-10)


def test_fused_image_of_more_rays_than_one_launch_holds(cuda_device):
    """
    A separable grid of 3 x 8.3e8 = 2.49e9 rays (> 2^31 - 1) through `image_rays`: the engine
    splits the outermost axis into launches.  The Newtonian has no dispersion, so every
    wavelength contributes the same hits: the counts are exactly three times a single one's.
    """
    system = configs.newtonian(num_field=900, num_pupil=32, num_pixel=64)
    edges = na.ScalarArray(np.array([400.0, 600.0]) * u.nm, "wavelength")
    one = system.image_rays(edges, wavelength=500 * u.nm, counts=True, **configs.PHYSICAL).counts.cpu().numpy()
    three = system.image_rays(
        edges, wavelength=na.ScalarArray(np.array([450.0, 500.0, 550.0]) * u.nm, "wavelength"), counts=True, **configs.PHYSICAL
    ).counts.cpu().numpy()
    assert 900 * 900 * 32 * 32 * 3 > 2**31 - 1
    assert one.sum() > 0 and np.array_equal(three, 3 * one)


def test_fused_image_in_strided_cta_order_matches_oracle(cuda_device):
    """
    More than 512 CTAs: fused image launches visit the ray grid in a strided CTA order
    (``TraceParams::cta_rows``: CTA b works on tile (b mod 512) * rows + b / 512, the last row is
    padded with CTAs that exit).  692 224 rays = 1352 tiles = 3 rows, 184 padding CTAs: every ray
    must still be traced and counted exactly once, on the fused path and in ``optk_bin``.
    """
    system = configs.newtonian(num_field=26, num_pupil=32, num_pixel=128)
    edges = na.ScalarArray(np.array([499.0, 501.0]) * u.nm, "wavelength")
    fused = system.image_rays(edges, counts=True, **configs.PHYSICAL)
    _, rays_in = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays_in)
    out = ora.propagate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
    assert out["px"].size == 26 * 26 * 32 * 32 > 512 * 512
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    want = orb.counts(local, edges.ndarray, ex, ey)
    got = fused.counts.cpu().numpy().reshape(want.shape)
    assert got.sum() == want.sum() > 0
    mism = got != want
    near = (orb.bin_margin(local["px"], ex) < 1e-9 * abs(ex[0])) | (orb.bin_margin(local["py"], ey) < 1e-9 * abs(ey[0]))
    assert mism.sum() <= 2 * near.sum()
    # the standalone kernel on the traced rays (> 512 CTAs of 256 rays as well)
    rays = system.rayfunction(on_device=True, **configs.PHYSICAL).outputs
    image, _ = system.sensor.collect(rays, wavelength=edges)
    assert np.allclose(fused.flux.cpu().numpy().reshape(image.outputs.ndarray.shape), image.outputs.ndarray, rtol=1e-12)
