"""
Parity of the on-device ray generator (``optk_trace_grid``) and of
``SequentialSystem.image`` against the oracle's restatement of
``optika/systems/_sequential.py:1002-1206``: generated rays, traced rays and the fused
detector image, on the same counter-based random stream.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import _engine, _grid, _lib
from oracle import raytrace as ora, binning as orb, grid as og

import configs
import parity

pytestmark = pytest.mark.gpu


def vertices(n=(3, 4, 5, 6, 7)):
    return [
        np.linspace(499e-6, 501e-6, n[0] + 1),
        np.linspace(-0.1, 0.1, n[1] + 1) * u.deg,
        np.linspace(-0.08, 0.1, n[2] + 1) * u.deg,
        np.linspace(-40, 40, n[3] + 1),
        np.linspace(-40, 38, n[4] + 1),
    ]


def device_dict(rays: _engine.DeviceRays) -> dict:
    out = {k: v.reshape(-1).cpu().numpy() for k, v in rays.fields.items()}
    out["unvignetted"] = rays.unvignetted.reshape(-1).cpu().numpy().astype(bool)
    return out


@pytest.fixture(scope="module")
def newtonian():
    return configs.newtonian(num_pixel=64)


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("at_infinity", [True, False])
def test_generated_rays_match_oracle(cuda_device, newtonian, jitter, at_infinity):
    v = vertices()
    if not at_infinity:
        v = [v[0], v[3], v[4], v[1], v[2]]
    rng = np.random.default_rng(0)
    n = [len(a) - 1 for a in v]
    ws, wp = rng.uniform(0.5, 2, n[:3]), rng.uniform(0.5, 2, n[3:])
    rot = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    frame = (rot, np.array([1.0, -2.0, 3.0])) if jitter else None
    grid = _grid.RayGrid(v, at_infinity, ws, wp, jitter=jitter, seed=1234567890123, frame=frame)
    got = _grid.trace_grid(newtonian._compiled_local, grid, surf_count=0)
    assert got.shape == dict(zip(_grid.AXES, n))
    want = og.input_rays(v, at_infinity, ws, wp, random=jitter, seed=1234567890123, frame=frame)
    parity.compare_states(device_dict(got), want)
    g = device_dict(got)
    assert np.abs(g["dx"] - want["dx"]).max() < 1e-15 and np.abs(g["dz"] - want["dz"]).max() < 1e-15


@pytest.mark.parametrize("n", [(1, 20, 20, 1, 1), (1, 1, 1, 1, 1), (2, 1, 3, 1, 300), (1, 1, 1, 70, 1)])
def test_single_cell_axes(cuda_device, newtonian, n):
    v = vertices(n)
    got = _grid.trace_grid(newtonian._compiled_local, _grid.RayGrid(v, seed=3), surf_count=0)
    parity.compare_states(device_dict(got), og.input_rays(v, seed=3))


def test_stream_is_independent_of_the_launch_split(cuda_device, newtonian):
    v = vertices()
    grid = _grid.RayGrid(v, seed=99)
    compiled = newtonian._compiled_local
    whole = device_dict(_grid.trace_grid(compiled, grid, surf_count=0))
    split = device_dict(_grid.trace_grid(compiled, grid, surf_count=0, max_launch=900))
    for k in whole:
        assert np.array_equal(whole[k], split[k]), k
    sub = grid.sub((1, 0, 2, 3, 0), (2, 4, 2, 2, 7))
    part = device_dict(_grid.trace_grid(compiled, sub, surf_count=0))
    full = whole["px"].reshape(3, 4, 5, 6, 7)[1:3, :, 2:4, 3:5, :].reshape(-1)
    assert np.array_equal(part["px"], full)
    # a different seed is a different sample
    other = device_dict(_grid.trace_grid(compiled, _grid.RayGrid(v, seed=100), surf_count=0))
    assert not np.array_equal(other["px"], whole["px"])


def test_traced_grid_matches_oracle_trace(cuda_device, newtonian):
    v = vertices((2, 5, 5, 24, 24))
    grid = _grid.RayGrid(v, seed=5)
    got, stats = _grid.trace_grid(newtonian._compiled_local, grid, stats=True)
    rays0 = og.input_rays(v, seed=5)
    want = ora.propagate_rays(newtonian.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(newtonian.sensor.transformation, want, inverse=True)
    parity.compare_states(device_dict(got), local)
    assert stats["n_rays"] == grid.size and stats["n_unvignetted"] == int(local["unvignetted"].sum())
    # accumulate: every surface, global coordinates for all but the (local) last one
    acc = _grid.trace_grid(newtonian._compiled, grid, accumulate=True, axis="surface")
    states = ora.accumulate_rays(newtonian.surfaces_all, rays0, extended=True)
    n_s = len(newtonian.surfaces_all)
    got_acc = {k: v.reshape(n_s, -1) for k, v in device_dict(acc).items()}
    parity.compare_states(got_acc, states)


def test_fused_grid_image_matches_oracle(cuda_device, newtonian):
    v = vertices((2, 6, 6, 32, 32))
    rng = np.random.default_rng(1)
    ws = rng.uniform(0.5, 2, (2, 6, 6))
    grid = _grid.RayGrid(v, weight_scene=ws, weight_pupil=rng.uniform(0.5, 2, (32, 32)), seed=17)
    ex, ey = newtonian.sensor.pixel_edges()
    ew = np.array([499e-6, 500e-6, 501e-6])
    compiled = newtonian._compiled_local
    image = _engine.DeviceImage.zeros(ew, ex, ey, cuda_device, moments=True, counts=True)
    _grid.trace_grid(compiled, grid, image=image, write_rays=False)
    rays0 = og.input_rays(v, weight_scene=ws, weight_pupil=grid.weight_pupil, seed=17)
    want = ora.propagate_rays(newtonian.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(newtonian.sensor.transformation, want, inverse=True)
    want_counts = orb.counts(local, ew, ex, ey)
    want_flux, want_direction, _ = orb.collect(local, ew, ex, ey)
    counts = image.counts.cpu().numpy()
    assert counts.sum() == want_counts.sum() > 0.5 * grid.size
    assert (counts != want_counts).sum() <= 4
    same = counts == want_counts
    flux = image.flux.cpu().numpy()
    assert np.allclose(flux[same], want_flux[same], rtol=1e-9, atol=1e-12)
    # sharded over four "ranks" and split into small launches: the same counts exactly
    sharded = _engine.DeviceImage.zeros(ew, ex, ey, cuda_device, moments=True, counts=True)
    for rank in range(4):
        _grid.trace_grid(compiled, grid.shard(rank, 4), image=sharded, write_rays=False, max_launch=5000)
    assert np.array_equal(sharded.counts.cpu().numpy(), counts)
    assert np.allclose(sharded.flux.cpu().numpy(), flux, rtol=1e-12, atol=1e-15)


def scene_for(system, num_field=(12, 10), num_wavelength=2, scale=1e9):
    field = na.Cartesian2dVectorLinearSpace(
        -0.1 * u.deg, 0.1 * u.deg, na.Cartesian2dVectorArray("field_x", "field_y"),
        na.Cartesian2dVectorArray(num_field[0] + 1, num_field[1] + 1),
    )
    w = na.linspace(499 * u.nm, 501 * u.nm, "wavelength", num_wavelength + 1)
    rng = np.random.default_rng(3)
    radiance = na.ScalarArray(
        rng.uniform(scale, 2 * scale, (num_wavelength,) + tuple(num_field)), ("wavelength", "field_x", "field_y")
    )
    return na.FunctionArray(
        inputs=optika.vectors.SpectralPositionalVectorArray(wavelength=w, position=field), outputs=radiance
    )


def oracle_image(system, scene, pupil_vertices, integrate, seed):
    w = scene.inputs.wavelength.ndarray
    f = scene.inputs.position
    v = [w, f.x.ndarray, f.y.ndarray, pupil_vertices[0], pupil_vertices[1]]
    aw, af, ap = og.cell_area(v, True, False)
    ws = scene.outputs.ndarray * aw[:, None, None] * af[None]
    rays0 = og.input_rays(v, weight_scene=ws, weight_pupil=ap, seed=seed)
    want = ora.propagate_rays(system.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(system.sensor.transformation, want, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    ew = np.array([w.min(), w.max()]) if integrate else w
    return orb.collect(local, ew, ex, ey)


@pytest.mark.parametrize("integrate", [True, False])
def test_image_of_a_scene_matches_oracle(cuda_device, newtonian, integrate):
    scene = scene_for(newtonian)
    pupil = na.Cartesian2dVectorLinearSpace(
        -40 * u.mm, 40 * u.mm, na.Cartesian2dVectorArray("pupil_x", "pupil_y"), 21
    )
    image = newtonian.image(scene, pupil=pupil, integrate=integrate, noise=False, normalized_pupil=False, seed=8)
    want_flux, _, _ = oracle_image(
        newtonian, scene, (pupil.x.ndarray, pupil.y.ndarray), integrate, seed=8
    )
    nw = 1 if integrate else 2
    assert image.outputs.shape == {"wavelength": nw, "detector_x": 64, "detector_y": 64}
    assert image.inputs.wavelength.shape == {"wavelength": nw + 1}
    got = image.outputs.ndarray * 1.0
    assert got.sum() > 0
    assert np.isclose(got.sum(), want_flux.sum(), rtol=1e-9)
    differ = ~np.isclose(got, want_flux, rtol=1e-9, atol=1e-9 * want_flux.max())
    assert differ.sum() <= 8  # rays within rounding of a pixel edge move between neighbours


def test_image_noise_and_default_pupil(cuda_device, newtonian):
    scene = scene_for(newtonian, num_field=(20, 20), num_wavelength=1, scale=1e12)
    clean = newtonian.image(scene, noise=False, seed=2)  # default: one normalised pupil cell
    again = newtonian.image(scene, noise=False, seed=2)
    assert np.array_equal(clean.outputs.ndarray, again.outputs.ndarray) or np.allclose(
        clean.outputs.ndarray, again.outputs.ndarray, rtol=1e-12
    )
    assert clean.outputs.ndarray.sum() > 0
    other = newtonian.image(scene, noise=False, seed=3)
    assert not np.allclose(other.outputs.ndarray, clean.outputs.ndarray)
    noisy = newtonian.image(scene, noise=True, seed=2)
    lam = clean.outputs.ndarray
    assert noisy.outputs.ndarray.shape == lam.shape
    z = (noisy.outputs.ndarray - lam)[lam > 50] / np.sqrt(lam[lam > 50])
    assert z.size > 20 and abs(z.mean()) < 0.5 and 0.5 < z.std() < 1.5


def test_image_over_a_configuration_axis(cuda_device):
    system = configs.misaligned_telescope(num_tilt=3, num_pixel=64)
    scene = scene_for(system, num_field=(6, 6), num_wavelength=1)
    pupil = na.Cartesian2dVectorLinearSpace(
        -30 * u.mm, 30 * u.mm, na.Cartesian2dVectorArray("pupil_x", "pupil_y"), 17
    )
    image = system.image(scene, pupil=pupil, noise=False, normalized_pupil=False, seed=4)
    (axis_config,) = system.shape
    assert image.outputs.shape[axis_config] == 3
    got = image.outputs.numpy((axis_config, "wavelength", "detector_x", "detector_y"))
    assert np.all(got.reshape(3, -1).sum(axis=1) > 0)
    # every configuration against the oracle, traced through that configuration's surfaces
    for c in range(3):
        surfaces = ora.select_config(system.surfaces_all, {axis_config: c})
        w, f = scene.inputs.wavelength.ndarray, scene.inputs.position
        v = [w, f.x.ndarray, f.y.ndarray, pupil.x.ndarray, pupil.y.ndarray]
        aw, af, ap = og.cell_area(v, True, False)
        rays0 = og.input_rays(v, weight_scene=scene.outputs.ndarray * aw[:, None, None] * af[None], weight_pupil=ap, seed=4)
        want = ora.propagate_rays(surfaces, rays0, extended=True)
        local = ora._rays_transform(surfaces[-1].transformation, want, inverse=True)
        ex, ey = system.sensor.pixel_edges()
        want_flux, _, _ = orb.collect(local, np.array([w.min(), w.max()]), ex, ey)
        assert np.isclose(got[c].sum(), want_flux.sum(), rtol=1e-9)
        assert (~np.isclose(got[c], want_flux, rtol=1e-9, atol=1e-9 * want_flux.max())).sum() <= 8


def test_grid_argument_errors(cuda_device, newtonian):
    grid = _grid.RayGrid(vertices())
    g = grid.struct(cuda_device)
    g.count[2] = 99
    rc = _lib.lib().optk_trace_grid(
        newtonian._compiled_local.handle, 0, g, None, 0, 0, 1, 0, 0, None, None, None, None
    )
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_more_rays_than_one_launch_holds(cuda_device, newtonian):
    """
    2.36e9 rays (> 2^31 - 1, the launch limit) through the fused generate + trace + bin path:
    the box is cut into launches; hit counts must equal the sum over two halves traced
    separately and the number of unvignetted rays the kernel reports.
    """
    n = (1, 1200, 1200, 40, 41)
    v = [
        np.linspace(499e-6, 501e-6, n[0] + 1), np.linspace(-0.1, 0.1, n[1] + 1) * u.deg,
        np.linspace(-0.1, 0.1, n[2] + 1) * u.deg, np.linspace(-40, 40, n[3] + 1), np.linspace(-40, 40, n[4] + 1),
    ]
    grid = _grid.RayGrid(v, seed=2)
    assert grid.size == 2_361_600_000 > 2**31 - 1
    compiled = newtonian._compiled_local
    ex, ey = newtonian.sensor.pixel_edges()
    ew = np.array([499e-6, 501e-6])
    whole = _engine.DeviceImage.zeros(ew, ex, ey, cuda_device, moments=False, counts=True)
    before = _grid.LAUNCHES
    _, stats = _grid.trace_grid(compiled, grid, image=whole, write_rays=False, stats=True)
    assert _grid.LAUNCHES - before >= 2
    assert stats["n_rays"] == grid.size
    halves = _engine.DeviceImage.zeros(ew, ex, ey, cuda_device, moments=False, counts=True)
    for begin, count in (((0, 0, 0, 0, 0), (1, 500, 1200, 40, 41)), ((0, 500, 0, 0, 0), (1, 700, 1200, 40, 41))):
        _grid.trace_grid(compiled, grid.sub(begin, count), image=halves, write_rays=False)
    a, b = whole.counts.cpu().numpy(), halves.counts.cpu().numpy()
    assert np.array_equal(a, b)
    assert 0 < a.sum() <= stats["n_unvignetted"]  # binned rays are unvignetted rays on the sensor
    assert a.sum() > 0.5 * stats["n_unvignetted"]


def test_image_with_normalized_field_and_pupil(cuda_device, newtonian):
    """
    Normalised scene and pupil vertices ([-1, 1], the reference's default units) are mapped to
    physical ones by the stop solver (``_denormalize_grid``, ``_sequential.py:748-789``) before
    the grid goes to the device; the oracle gets the same physical vertices.
    """
    nf, npup = 8, 12
    field = na.Cartesian2dVectorLinearSpace(-0.25, 0.25, na.Cartesian2dVectorArray("field_x", "field_y"), nf + 1)
    w = na.linspace(499 * u.nm, 501 * u.nm, "wavelength", 2)
    rng = np.random.default_rng(9)
    scene = na.FunctionArray(
        inputs=optika.vectors.SpectralPositionalVectorArray(wavelength=w, position=field),
        outputs=na.ScalarArray(rng.uniform(1e9, 2e9, (1, nf, nf)), ("wavelength", "field_x", "field_y")),
    )
    pupil = na.Cartesian2dVectorLinearSpace(-1, 1, na.Cartesian2dVectorArray("pupil_x", "pupil_y"), npup + 1)
    image = newtonian.image(
        scene, pupil=pupil, noise=False, normalized_field=True, normalized_pupil=True, seed=12
    )
    physical = newtonian.denormalize(
        optika.vectors.ObjectVectorArray(wavelength=w, field=field, pupil=pupil), True, True
    )
    vert = lambda a, ax: newtonian._separable(a, ax, {}, (), ax)  # noqa: E731
    v = [
        w.ndarray, vert(physical.field.x, "field_x"), vert(physical.field.y, "field_y"),
        vert(physical.pupil.x, "pupil_x"), vert(physical.pupil.y, "pupil_y"),
    ]
    assert np.isclose(v[3][0], -v[3][-1]) and 39 < v[3][-1] < 42  # the 40 mm primary is the pupil stop
    aw, af, ap = og.cell_area(v, True, False)
    rays0 = og.input_rays(v, weight_scene=scene.outputs.ndarray * aw[:, None, None] * af[None], weight_pupil=ap, seed=12)
    out = ora.propagate_rays(newtonian.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(newtonian.sensor.transformation, out, inverse=True)
    ex, ey = newtonian.sensor.pixel_edges()
    want, _, _ = orb.collect(local, np.array([w.ndarray.min(), w.ndarray.max()]), ex, ey)
    got = image.outputs.ndarray
    assert want.sum() > 0 and np.isclose(got.sum(), want.sum(), rtol=1e-9)
    assert (~np.isclose(got, want, rtol=1e-9, atol=1e-9 * want.max())).sum() <= 8


def polar_pupil(num_r=6, num_phi=12):
    r = na.linspace(5 * u.mm, 40 * u.mm, "pupil_r", num_r + 1)
    phi = na.linspace(0, 2 * np.pi, "pupil_phi", num_phi + 1)
    return na.Cartesian2dVectorArray(r * np.cos(phi), r * np.sin(phi))


@pytest.mark.parametrize("jitter", [False, True])
def test_curvilinear_grid_rays_match_oracle(cuda_device, newtonian, jitter):
    """Field and pupil given as 2-D vertex arrays (a sheared field grid, a polar pupil grid)."""
    fx, fy = np.meshgrid(np.linspace(-0.1, 0.1, 5) * u.deg, np.linspace(-0.08, 0.1, 4) * u.deg, indexing="ij")
    fx = fx + 0.2 * fy  # sheared: not separable
    p = polar_pupil()
    px, py = p.x.numpy(("pupil_r", "pupil_phi")), p.y.numpy(("pupil_r", "pupil_phi"))
    v = [np.linspace(499e-6, 501e-6, 3), fx, fy, px, py]
    grid = _grid.RayGrid(v, jitter=jitter, seed=31)
    assert grid.n == (2, 4, 3, 6, 12) and grid.field_2d and grid.pupil_2d
    got = _grid.trace_grid(newtonian._compiled_local, grid, surf_count=0)
    want = og.input_rays(v, random=jitter, seed=31)
    parity.compare_states(device_dict(got), want)
    # sub-boxes draw the same samples
    sub = grid.sub((1, 1, 0, 2, 3), (1, 2, 3, 3, 5))
    part = device_dict(_grid.trace_grid(newtonian._compiled_local, sub, surf_count=0))
    full = device_dict(got)["px"].reshape(grid.n)[1:2, 1:3, :, 2:5, 3:8].reshape(-1)
    assert np.array_equal(part["px"], full)


def test_image_with_a_polar_pupil_grid(cuda_device, newtonian):
    scene = scene_for(newtonian, num_field=(6, 6), num_wavelength=1)
    pupil = polar_pupil(num_r=8, num_phi=24)
    image = newtonian.image(
        scene, pupil=pupil, axis_pupil=("pupil_r", "pupil_phi"), noise=False, normalized_pupil=False, seed=5
    )
    w, f = scene.inputs.wavelength.ndarray, scene.inputs.position
    v = [w, f.x.ndarray, f.y.ndarray, pupil.x.numpy(("pupil_r", "pupil_phi")), pupil.y.numpy(("pupil_r", "pupil_phi"))]
    aw, af, ap = og.cell_area(v, True, False)
    assert np.isclose(ap.sum(), np.pi * (40**2 - 5**2) * np.sin(2 * np.pi / 24) / (2 * np.pi / 24), rtol=1e-12)
    rays0 = og.input_rays(v, weight_scene=scene.outputs.ndarray * aw[:, None, None] * af[None], weight_pupil=ap, seed=5)
    out = ora.propagate_rays(newtonian.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(newtonian.sensor.transformation, out, inverse=True)
    ex, ey = newtonian.sensor.pixel_edges()
    want, _, _ = orb.collect(local, np.array([w.min(), w.max()]), ex, ey)
    got = image.outputs.ndarray
    assert want.sum() > 0 and np.isclose(got.sum(), want.sum(), rtol=1e-9)
    assert (~np.isclose(got, want, rtol=1e-9, atol=1e-9 * want.max())).sum() <= 8


# ---------------------------------------------------------------------------
# chromatic grids: field / pupil vertices that depend on the wavelength (optk_grid_t.chromatic)
# ---------------------------------------------------------------------------
def chromatic_vertices(n=(3, 4, 5, 6, 7), axes=(1, 4)):
    v = vertices(n)
    stretch = np.linspace(1.0, 1.3, n[0] + 1)
    for a in axes:
        v[a] = stretch[:, None] * v[a][None, :] + (0.02 * stretch[:, None] if a > 2 else 0.0)
    return v


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("axes", [(1,), (1, 2), (3, 4), (1, 4), (1, 2, 3, 4)])
def test_chromatic_generated_rays_match_oracle(cuda_device, newtonian, jitter, axes):
    v = chromatic_vertices(axes=axes)
    rng = np.random.default_rng(0)
    n = (3, 4, 5, 6, 7)
    ws = rng.uniform(0.5, 2, n[:3])
    wp = rng.uniform(0.5, 2, (n[0],) + n[3:]) if set(axes) & {3, 4} else rng.uniform(0.5, 2, n[3:])
    grid = _grid.RayGrid(v, chromatic=axes, weight_scene=ws, weight_pupil=wp, jitter=jitter, seed=77)
    assert grid.n == n
    got = _grid.trace_grid(newtonian._compiled_local, grid, surf_count=0)
    want = og.input_rays(v, True, ws, wp, random=jitter, seed=77, chromatic=axes)
    parity.compare_states(device_dict(got), want)
    # a sub-box sees the same stream
    sub = grid.sub((1, 0, 2, 3, 0), (2, 4, 2, 2, 7))
    part = device_dict(_grid.trace_grid(newtonian._compiled_local, sub, surf_count=0))
    full = device_dict(got)["dx"].reshape(n)[1:3, :, 2:4, 3:5, :].reshape(-1)
    assert np.array_equal(part["dx"], full)


def test_chromatic_fused_image_matches_oracle_and_the_specialised_kernel(cuda_device, newtonian):
    from optika_b200 import _lib

    v = chromatic_vertices((2, 6, 6, 32, 32), axes=(1, 2))
    grid = _grid.RayGrid(v, chromatic=(1, 2), seed=17)
    ex, ey = newtonian.sensor.pixel_edges()
    ew = np.array([499e-6, 500e-6, 501e-6])
    compiled = newtonian._compiled_local
    rays0 = og.input_rays(v, seed=17, chromatic=(1, 2))
    want = ora.propagate_rays(newtonian.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(newtonian.sensor.transformation, want, inverse=True)
    want_counts = orb.counts(local, ew, ex, ey)
    images = []
    for mode in (0, 1):  # table-driven kernels, then a kernel compiled for this walk and this kind of grid
        _lib.check(_lib.lib().optk_jit_mode(mode))
        try:
            image = _engine.DeviceImage.zeros(ew, ex, ey, cuda_device, moments=True, counts=True)
            _grid.trace_grid(compiled, grid, image=image, write_rays=False)
            images.append(image.counts.cpu().numpy())
        finally:
            _lib.check(_lib.lib().optk_jit_mode(-1))
    assert images[0].sum() == want_counts.sum() > 0.3 * grid.size
    assert (images[0] != want_counts).sum() <= 4
    assert np.array_equal(images[0], images[1])


def test_image_with_a_chromatic_stop_solution(cuda_device):
    """
    ``image(..., normalized_field=True)`` of a spectrograph: the field of view comes from a stop solution
    per WAVELENGTH (``_sequential.py:748-789``), so the denormalised field vertices depend on the wavelength.
    The image must equal the oracle's on the rays of the very grid ``ray_grids`` builds.
    """
    system = configs.spherical_grating(num_field=4, num_pupil=8, num_wavelength=3, num_pixel=128)
    wavelength = na.linspace(38 * u.nm, 42 * u.nm, axis="wavelength", num=4)
    field = na.Cartesian2dVectorLinearSpace(-0.6, 0.6, axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=7)
    pupil = na.Cartesian2dVectorLinearSpace(-0.9, 0.9, axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=17)
    axes = ("wavelength", ("field_x", "field_y"), ("pupil_x", "pupil_y"))
    (grid,) = system.ray_grids(1.0, wavelength, field, pupil, *axes, normalized_field=True, normalized_pupil=True, seed=4)
    assert grid.chromatic and set(grid.chromatic) <= {1, 2, 3, 4}  # the sensor is the field stop: its image moves with wavelength
    w_edges = np.array([38 * u.nm, 42 * u.nm])
    planes = system.collect_grids([grid], w_edges, device=cuda_device, counts=True)
    rays0 = og.input_rays(grid.vertices, True, grid.weight_scene, grid.weight_pupil, seed=4, chromatic=grid.chromatic)
    want = ora.propagate_rays(system.surfaces_all, rays0, extended=True)
    local = ora._rays_transform(system.sensor.transformation, want, inverse=True)
    ex, ey = system.sensor.pixel_edges()
    want_counts = orb.counts(local, w_edges, ex, ey)
    assert planes["counts"].sum() == want_counts.sum() > 0.5 * grid.size
    assert (planes["counts"].reshape(want_counts.shape) != want_counts).sum() <= 4
    want_flux, _, _ = orb.collect(local, w_edges, ex, ey)
    assert np.isclose(planes["flux"].sum(), want_flux.sum(), rtol=1e-9)
