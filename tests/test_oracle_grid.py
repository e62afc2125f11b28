"""
Pins of the oracle's input-grid restatement (``oracle/grid.py``): the Philox4x32-10
known-answer vectors of Random123, the stratified-sample contract of
``cell_centers(random=True)`` (``optika/systems/_sequential.py:1066-1069``), cell areas
against closed forms (``optika/vectors/_vectors_object.py:98-133``), and the host side of
the product (``RayGrid``, ``SequentialSystem.ray_grids``) against the oracle.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import _grid
from oracle import grid as og

import configs


def words(counter, key):
    return [int(v[0]) for v in og.philox4x32_10([np.array([c]) for c in counter], key)]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert words((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert words((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert words((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1,
    ]


def test_jitter_is_uniform_and_independent():
    t = og.jitter(np.arange(200_000), seed=7)
    assert t.shape == (5, 200_000)
    assert t.min() > 0 and t.max() < 1
    assert np.allclose(t.mean(axis=1), 0.5, atol=5e-3)
    assert np.allclose(t.var(axis=1), 1 / 12, atol=2e-3)
    c = np.corrcoef(t)
    assert np.abs(c - np.eye(5)).max() < 0.01
    assert not np.array_equal(t, og.jitter(np.arange(200_000), seed=8))
    # 64-bit cell indices: the high word reaches the counter
    assert not np.array_equal(og.jitter(np.array([5]), 0), og.jitter(np.array([5 + 2**32]), 0))


def vertices(n=(3, 4, 5, 6, 7)):
    return [
        np.linspace(1e-5, 2e-5, n[0] + 1),
        np.linspace(-0.01, 0.02, n[1] + 1),
        np.linspace(-0.015, 0.01, n[2] + 1) ** 3 * 1e4,  # non-uniform
        np.linspace(-30, 40, n[3] + 1),
        np.linspace(-20, 20, n[4] + 1),
    ]


def test_one_sample_per_cell_inside_the_cell():
    v = vertices()
    samples, idx = og.cell_samples(v, random=True, seed=3)
    for a in range(5):
        lo, hi = np.minimum(v[a][idx[a]], v[a][idx[a] + 1]), np.maximum(v[a][idx[a]], v[a][idx[a] + 1])
        assert samples[a].shape == (3, 4, 5, 6, 7)
        assert np.all((samples[a] > lo) & (samples[a] < hi))
    centres, _ = og.cell_samples(v, random=False)
    assert np.array_equal(centres[3][0, 0, 0, :, 0], (v[3][1:] + v[3][:-1]) / 2)


def test_stream_does_not_depend_on_the_sub_box():
    v = vertices()
    full, _ = og.cell_samples(v, seed=11)
    part, _ = og.cell_samples(v, begin=(1, 0, 2, 3, 0), count=(2, 4, 2, 2, 7), seed=11)
    for a in range(5):
        assert np.array_equal(part[a], full[a][1:3, :, 2:4, 3:5, :])


def test_input_rays_follow_the_object_location():
    v = vertices()
    r = og.input_rays(v, at_infinity=True, random=False)
    assert np.all(r["pz"] == 0) and np.all(r["index_refraction"] == 1) and r["unvignetted"].all()
    assert np.allclose(r["dx"] ** 2 + r["dy"] ** 2 + r["dz"] ** 2, 1, atol=1e-15)
    # optika.direction: positive field angle about y tips the ray towards -x (optika/_util_test.py:23-32)
    one = og.direction(np.array(0.1), np.array(0.0))
    assert np.isclose(one[0], -np.sin(0.1)) and one[1] == 0 and np.isclose(one[2], np.cos(0.1))
    f = og.input_rays(v, at_infinity=False, random=False)
    assert np.array_equal(np.unique(f["px"]), np.unique((v[1][1:] + v[1][:-1]) / 2))
    rot = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    g = og.input_rays(v, random=False, frame=(rot, np.array([1.0, 2.0, 3.0])))
    assert np.allclose(g["px"], -r["py"] + 1) and np.allclose(g["py"], r["px"] + 2) and np.allclose(g["pz"], 3)
    assert np.allclose(g["dx"], -r["dy"]) and np.allclose(g["dy"], r["dx"])


def test_cell_areas_closed_forms():
    # the whole sphere
    ax, ay = np.meshgrid(np.linspace(-np.pi, np.pi, 37), np.linspace(-np.pi / 2, np.pi / 2, 19), indexing="ij")
    a = og.solid_angle_cell(ax, ay)
    assert np.isclose(a.sum(), 4 * np.pi, rtol=1e-12) and a.min() > 0
    # one octant (a spherical triangle with three right angles), as a degenerate quadrilateral
    ox, oy = np.meshgrid(np.array([0.0, np.pi / 2]), np.array([0.0, np.pi / 2]), indexing="ij")
    assert np.isclose(abs(og.solid_angle_cell(ox, oy)[0, 0]), np.pi / 2, rtol=1e-14)
    # small cells: d(azimuth) d(elevation) cos(elevation)
    sx, sy = np.meshgrid(0.3 + np.array([0.0, 1e-4]), 0.2 + np.array([0.0, 2e-4]), indexing="ij")
    assert np.isclose(abs(og.solid_angle_cell(sx, sy)[0, 0]), 2e-8 * np.cos(0.2001), rtol=1e-6)
    # planar quadrilaterals
    x, y = np.meshgrid(np.array([0.0, 2.0, 5.0]), np.array([1.0, 2.0, 4.0, 8.0]), indexing="ij")
    assert np.array_equal(og.volume_cell_2d(x, y), np.outer([2.0, 3.0], [1.0, 2.0, 4.0]))
    assert np.array_equal(og.volume_cell_1d([1.0, 3.0, 2.0]), [2.0, -1.0])
    w, f, p = og.cell_area(vertices(), True, False)
    assert w.shape == (3,) and f.shape == (4, 5) and p.shape == (6, 7) and f.min() > 0 and p.min() > 0


def test_named_cell_area_matches_oracle():
    v = vertices()
    grid = optika.vectors.ObjectVectorArray(
        wavelength=na.ScalarArray(v[0], "w"),
        field=na.Cartesian2dVectorArray(na.ScalarArray(v[1], "fx"), na.ScalarArray(v[2], "fy")),
        pupil=na.Cartesian2dVectorArray(na.ScalarArray(v[3], "px"), na.ScalarArray(v[4], "py")),
    )
    area = grid.cell_area("w", ("fx", "fy"), ("px", "py"))
    w, f, p = og.cell_area(v, True, False)
    want = w[:, None, None, None, None] * f[None, :, :, None, None] * p[None, None, None]
    assert np.allclose(area.numpy(("w", "fx", "fy", "px", "py")), want, rtol=1e-12)
    with pytest.raises(ValueError):
        grid.cell_area("nope", ("fx", "fy"), ("px", "py"))


def test_boxes_cover_the_grid_exactly():
    count = (3, 4, 5, 6, 7)
    seen = np.zeros(count, dtype=int)
    for begin, c in _grid._boxes((0,) * 5, count, limit=100):
        assert np.prod(c) <= 100
        seen[tuple(slice(b, b + n) for b, n in zip(begin, c))] += 1
    assert np.all(seen == 1)
    assert list(_grid._boxes((1, 0, 0, 0, 0), count, 10**6)) == [((1, 0, 0, 0, 0), count)]


def test_ray_grid_shards_partition_the_box():
    g = _grid.RayGrid(vertices=vertices())
    assert g.n == (3, 4, 5, 6, 7) and g.size == 2520 and g.shape["pupil_x"] == 6
    parts = [g.shard(r, 4, axis=3) for r in range(4)]
    assert sum(p.size for p in parts) == g.size
    assert [p.begin[3] for p in parts] == [0, 2, 4, 5] and [p.count[3] for p in parts] == [2, 2, 1, 1]
    with pytest.raises(ValueError):
        _grid.RayGrid(vertices=vertices()[:4])


class _NoDevice:
    """ray_grids only needs the configuration shape of the lowered system."""


def test_ray_grids_carry_radiance_times_area():
    system = configs.newtonian(num_pixel=64)
    w = na.linspace(499 * u.nm, 501 * u.nm, "wavelength", 3)
    field = na.Cartesian2dVectorLinearSpace(
        -0.1 * u.deg, 0.1 * u.deg, na.Cartesian2dVectorArray("field_x", "field_y"), na.Cartesian2dVectorArray(5, 4)
    )
    pupil = na.Cartesian2dVectorLinearSpace(
        -40 * u.mm, 40 * u.mm, na.Cartesian2dVectorArray("pupil_x", "pupil_y"), 7
    )
    radiance = na.ScalarArray(np.arange(1.0, 25.0).reshape(2, 4, 3), ("wavelength", "field_x", "field_y"))
    try:
        grids = system.ray_grids(
            radiance, w, field, pupil, "wavelength", ("field_x", "field_y"), ("pupil_x", "pupil_y"),
            normalized_field=False, normalized_pupil=False, seed=5,
        )
    except Exception as e:  # the lowered table needs liboptk.so, not a GPU
        pytest.skip(f"liboptk.so unavailable: {e}")
    (g,) = grids
    assert g.n == (2, 4, 3, 6, 6) and g.at_infinity and g.jitter and g.seed == 5 and g.frame is None
    aw, af, ap = og.cell_area(g.vertices, True, False)
    assert np.allclose(g.weight_scene, radiance.ndarray * aw[:, None, None] * af[None], rtol=1e-12)
    assert np.allclose(g.weight_pupil, ap, rtol=1e-12)
    # field vertices that depend on the wavelength (a chromatic stop solution): one row per wavelength vertex
    scale = na.linspace(1, 2, "wavelength", 3)
    chromatic_field = na.Cartesian2dVectorArray(field.x * scale, field.y)
    (gc,) = system.ray_grids(
        radiance, w, chromatic_field, pupil, "wavelength", ("field_x", "field_y"), ("pupil_x", "pupil_y"), False, False
    )
    assert gc.chromatic == (1,) and gc.n == g.n and gc.vertices[1].shape == (3, 5) and gc.vertices[2].ndim == 1
    assert np.allclose(gc.vertices[1], scale.ndarray[:, None] * g.vertices[1][None, :], rtol=1e-15)
    assert gc.angular_cells() == [None, None] and gc.struct is not None
    # cell areas: the solid angle of every (wavelength vertex, field cell), averaged over the two vertices of the cell
    per_vertex = np.stack([og.cell_area([g.vertices[0], row, g.vertices[2], g.vertices[3], g.vertices[4]], True, False)[1]
                           for row in gc.vertices[1]])
    area_f = 0.5 * (per_vertex[1:] + per_vertex[:-1])
    assert np.allclose(gc.weight_scene, radiance.ndarray * aw[:, None, None] * area_f, rtol=1e-12)
    assert np.allclose(gc.weight_pupil, ap, rtol=1e-12)
    # vertices that vary along anything else are still refused
    with pytest.raises(NotImplementedError):
        bad = na.Cartesian2dVectorArray(field.x * na.linspace(1, 2, "other", 3), field.y)
        system.ray_grids(
            radiance, w, bad, pupil, "wavelength", ("field_x", "field_y"), ("pupil_x", "pupil_y"), False, False
        )


def test_chromatic_cell_samples_are_bilinear_in_wavelength_and_axis():
    w = np.array([1.0, 2.0, 4.0])
    fx = np.array([[0.0, 1.0, 2.0], [0.0, 2.0, 4.0], [0.0, 4.0, 8.0]])  # [wavelength vertex][field_x vertex]
    other = np.array([0.0, 1.0])
    vertices = [w, fx, other, other, other]
    (sw, sfx, sfy, spx, spy), idx = og.cell_samples(vertices, random=False, chromatic=(1,))
    assert sfx.shape == (2, 2, 1, 1, 1)
    # cell centres: the mean of the four corners
    assert np.allclose(sfx[:, :, 0, 0, 0], [[0.75, 2.25], [1.5, 4.5]])
    (sw, sfx, *_), idx = og.cell_samples(vertices, random=True, seed=3, chromatic=(1,))
    t = og.jitter(((idx[0] * 2 + idx[1]) * 1 + 0).astype(np.uint64), 3)
    lo = fx[idx[0], idx[1]] + t[0] * (fx[idx[0] + 1, idx[1]] - fx[idx[0], idx[1]])
    hi = fx[idx[0], idx[1] + 1] + t[0] * (fx[idx[0] + 1, idx[1] + 1] - fx[idx[0], idx[1] + 1])
    assert np.allclose(sfx, lo + t[1] * (hi - lo), rtol=1e-15)
    rays = og.input_rays(vertices, weight_pupil=np.arange(2.0).reshape(2, 1, 1) + 1, chromatic=(1,), seed=3)
    assert np.allclose(rays["intensity"].reshape(2, 2), [[1, 1], [2, 2]])
