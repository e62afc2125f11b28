"""Host-side logic: named arrays, transformations, lowering, flattening, sharding (CPU only)."""

import ctypes as C

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import transformations as tf
from optika_b200 import _lowering, _lib, _engine, distributed
from optika_b200.materials._multilayers import flatten_layers
from oracle import raytrace as ora

import configs


def test_named_broadcasting_by_axis_name():
    a = na.ScalarArray(np.arange(3.0), "x")
    b = na.ScalarArray(np.arange(4.0) * 10, "y")
    c = a + b
    assert c.shape == {"x": 3, "y": 4}
    assert c.ndarray[2, 3] == 32
    d = np.sin(a) * 2 - b
    assert d.shape == {"x": 3, "y": 4}
    assert (b + a).shape == {"y": 4, "x": 3}
    assert np.allclose(np.asarray((b + a).numpy(("x", "y"))), c.ndarray)
    assert c[dict(x=1)].shape == {"y": 4}
    assert c[dict(y=slice(0, 2))].shape == {"x": 3, "y": 2}
    assert c.max("x").shape == {"y": 4}
    with pytest.raises(ValueError):
        na.ScalarArray(np.arange(3.0), "x") + na.ScalarArray(np.arange(4.0), "x")


def test_linspace_centers_and_vector_space():
    a = na.linspace(-1, 1, axis="p", num=4, centers=True)
    assert np.allclose(a.ndarray, [-0.75, -0.25, 0.25, 0.75])
    v = na.Cartesian2dVectorLinearSpace(-1, 1, axis=na.Cartesian2dVectorArray("px", "py"), num=5)
    assert v.x.axes == ("px",) and v.y.axes == ("py",)
    assert v.shape == {"px": 5, "py": 5}
    w = na.Cartesian3dVectorArray(1.0, 2.0, 2.0)
    assert w.length == 3.0
    assert np.isclose((w.normalized @ w.normalized), 1.0)
    z = na.Cartesian3dVectorArray(1, 0, 0).cross(na.Cartesian3dVectorArray(0, 1, 0))
    assert (z.x, z.y, z.z) == (0, 0, 1)


TRANSFORMS = [
    tf.Cartesian3dTranslation(x=5.0, y=-1.0, z=2.0),
    tf.Cartesian3dRotationX(0.3),
    tf.Cartesian3dRotationY(-0.7),
    tf.Cartesian3dRotationZ(53 * u.deg),
    tf.TransformationList(
        [
            tf.Cartesian3dTranslation(x=5.0),
            tf.Cartesian3dRotationZ(53 * u.deg),
            tf.Cartesian3dTranslation(x=6.0),
            tf.Cartesian3dRotationY(0.2),
        ]
    ),
]


@pytest.mark.parametrize("t", TRANSFORMS)
def test_composed_affine_matches_sequential_primitives(t):
    """The product composes one affine; the oracle applies the primitives one by one."""
    rng = np.random.default_rng(0)
    x, y, z = rng.normal(size=(3, 50))
    r, v = t.affine.numpy({})
    got = r @ np.stack([x, y, z]) + v[:, None]
    want = np.stack(ora.transform_forward(t, x, y, z))
    assert np.allclose(got, want, atol=1e-13)
    ri, vi = t.inverse.affine.numpy({})
    got_inv = ri @ np.stack([x, y, z]) + vi[:, None]
    want_inv = np.stack(ora.transform_inverse(t, x, y, z))
    assert np.allclose(got_inv, want_inv, atol=1e-13)
    # direction: linear part only
    want_d = np.stack(ora.transform_forward(t, x, y, z, is_direction=True))
    assert np.allclose(r @ np.stack([x, y, z]), want_d, atol=1e-13)


def test_transformation_with_configuration_axis():
    t = tf.TransformationList(
        [tf.Cartesian3dRotationX(na.linspace(-0.1, 0.1, "tilt", 3)), tf.Cartesian3dTranslation(z=200.0)]
    )
    assert t.shape == {"tilt": 3}
    r, v = t.affine.numpy({"tilt": 3})
    assert r.shape == (3, 3, 3) and v.shape == (3, 3)
    assert np.allclose(r[1], np.eye(3))
    assert np.allclose(v[:, 2], 200.0)


def test_lowering_newtonian_table():
    system = configs.newtonian()
    table, shape_ = _lowering.lower_system(system.surfaces_all)
    assert shape_ == {} and len(table) == 6
    kinds = [(s.sag_kind, s.material_kind, s.aperture_kind) for s in table]
    assert kinds[0] == (_lib.SAG_FLAT, _lib.MAT_VACUUM, _lib.APERTURE_NONE)
    assert kinds[2] == (_lib.SAG_FLAT, _lib.MAT_VACUUM, _lib.APERTURE_RECTANGULAR)
    assert kinds[3] == (_lib.SAG_PARABOLIC, _lib.MAT_MIRROR, _lib.APERTURE_RECTANGULAR)
    assert table[2].flags & _lib.F_APERTURE_INVERTED
    assert table[3].sag[0] == -200.0
    assert table[3].transform.t[2] == 200.0
    assert not (table[0].flags & _lib.F_TRANSFORM) and (table[5].flags & _lib.F_TRANSFORM)
    # fold mirror: RotationY(135 deg) then translate z = 50
    r = np.array(table[4].transform.r[:]).reshape(3, 3)
    assert np.allclose(r @ [0, 0, 1], [np.sin(np.deg2rad(135)), 0, np.cos(np.deg2rad(135))])
    assert list(table[4].transform.t[:]) == [0.0, 0.0, 50.0]


def test_lowering_configuration_axes():
    system = configs.misaligned_telescope(num_tilt=4)
    table, shape_ = _lowering.lower_system(system.surfaces_all)
    assert shape_ == {"misalign": 4} and len(table) == 24
    tilts = [np.arcsin(table[c * 6 + 3].transform.r[7]) for c in range(4)]
    assert np.allclose(tilts, np.linspace(-30, 30, 4) * u.arcsec)


def test_lowering_rulings_and_polygon():
    system = configs.toroidal_vls()
    table, _ = _lowering.lower_system(system.surfaces_all)
    stop, grating = table[1], table[2]
    assert stop.aperture_kind == _lib.APERTURE_POLYGON and stop.n_vertices == 8
    assert np.isclose(stop.vertices_x[0], 20.0) and abs(stop.vertices_y[0]) < 1e-12
    assert grating.sag_kind == _lib.SAG_TOROIDAL and grating.ruling_kind == _lib.RULING_POLYNOMIAL
    assert grating.n_coeff == 3 and list(grating.ruling_power[:3]) == [0, 1, 2]
    assert np.isclose(grating.ruling_coeff[0], 1 / 2400)


def test_lowering_rejects_unsupported():
    class Strange(optika.materials.AbstractMaterial):
        pass

    with pytest.raises(NotImplementedError):
        _lowering.lower_system([optika.surfaces.Surface(material=Strange())])


def test_merge_axes():
    dims, strides = _engine._merge_axes([4, 1, 5, 6], [[30, 0, 6, 1], [0, 0, 0, 0]])
    assert dims == [120] and strides == [[1], [0]]
    dims, strides = _engine._merge_axes([4, 5, 6], [[0, 6, 1], [1, 0, 0]])
    assert dims == [4, 30] and strides == [[0, 1], [1, 0]]
    dims, strides = _engine._merge_axes([3, 4], [[1, 3]])
    assert dims == [3, 4]


def test_flatten_layers_segments():
    M = optika.materials
    a, b, c = M.Layer("Si", thickness=1e-6), M.Layer("Mo", thickness=2e-6), M.Layer("SiO2", thickness=3e-6)
    flat, seg = flatten_layers([c, M.PeriodicLayerSequence([a, b], num_periods=30), c])
    assert [x.chemical for x in flat] == ["SiO2", "Si", "Mo", "SiO2"]
    assert seg == [(0, 1, 1), (1, 2, 30), (3, 1, 1)]
    flat, seg = flatten_layers(M.LayerSequence([a, b, M.LayerSequence([c])]))
    assert seg == [(0, 3, 1)]
    flat, seg = flatten_layers(None)
    assert flat == [] and seg == []


def test_chemical_table_interpolation():
    si = optika.chemicals.Chemical("Si")
    w = na.ScalarArray(np.array([100.0, 150.0]) * u.AA, "wavelength")
    n = si.n(w)
    assert n.shape == {"wavelength": 2} and np.iscomplexobj(n.ndarray)
    from oracle import multilayer as orm
    import pathlib

    table = orm.load_nk(pathlib.Path(optika.chemicals._PATH_BUNDLED) / "Si.nk")
    assert np.allclose(n.ndarray, orm.interp_nk(np.array([100.0, 150.0]), table))
    assert optika.chemicals.Chemical("Si", is_amorphous=True, table="x").file_nk == "a-Si_x.nk"


def test_slab_partition_is_exact():
    for n in (0, 1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            covered = []
            for rank in range(world):
                s = distributed.slab(n, rank, world)
                covered += list(range(n))[s]
            assert covered == list(range(n))
    with pytest.raises(ValueError):
        distributed.slab(10, 2, 2)


def test_shard_grid_by_named_axis():
    system = configs.spherical_grating(num_field=4, num_pupil=10, num_wavelength=3)
    parts = [distributed.shard_grid(system.grid_input, "pupil_x", r, 4) for r in range(4)]
    assert [p.pupil.x.shape["pupil_x"] for p in parts] == [3, 3, 2, 2]
    assert all(p.pupil.y.shape == {"pupil_y": 10} for p in parts)
    joined = np.concatenate([p.pupil.x.ndarray for p in parts])
    assert np.array_equal(joined, system.grid_input.pupil.x.ndarray)


def test_input_rays_match_reference_construction():
    """_calc_rayfunction_input (optika/systems/_sequential.py:791-828): object at infinity."""
    system = configs.newtonian(num_field=3, num_pupil=4)
    _, rays = system._input(None, None, None, None, False, False)
    assert rays.shape == {"pupil_x": 4, "pupil_y": 4, "field_y": 3, "field_x": 3}
    fx = system.grid_input.field.x.ndarray
    fy = system.grid_input.field.y.ndarray
    want = ora.direction(fx[:, None], fy[None, :])
    assert np.allclose(rays.direction.x.numpy(("field_x", "field_y")), want[0])
    assert np.allclose(rays.direction.z.numpy(("field_x", "field_y")), want[2])
    assert float(rays.position.z) == 0.0
    assert system._ray_axes_order == ["field_x", "field_y", "pupil_x", "pupil_y"]


def test_device_groups_struct_addresses_whole_groups():
    """``DeviceGroups.struct``: plane pointers of one configuration, offset to the first group of a launch."""
    import torch

    groups = _engine.DeviceGroups.zeros(n_config=3, n_groups=10, group_size=64, device=torch.device("cpu"))
    im = groups.struct(0)
    assert (im.n_wavelength, im.n_x, im.n_y, im.group_size) == (1, 10, 1, 64)
    assert im.flux == groups.flux.data_ptr() and im.counts == groups.counts.data_ptr()
    im = groups.struct(2, first_ray=4 * 64)  # third configuration, launch starting at the fifth group
    assert im.n_x == 6
    assert im.moment_imag == groups.moment_imag.data_ptr() + 8 * (2 * 10 + 4)
    with pytest.raises(ValueError):
        groups.struct(0, first_ray=65)  # not a group boundary


def test_reductions_have_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    system = configs.newtonian(num_field=2, num_pupil=4)
    with pytest.raises(RuntimeError):
        system.pupil_moments(**configs.PHYSICAL)


def test_solid_angles_of_direction_grids_blocked_and_threaded():
    """
    ``Cartesian3dVectorArray.solid_angle_cell`` works on blocks of rows with a few threads (a 4096 x 4096 field grid
    has 1.7e7 cells): same values as the oracle's plain formula, with leading axes, either axis order, vertices that
    are not unit vectors, and a grid large enough to be split.  (The triple product carries an absolute rounding
    error of ~1e-16 sr in either evaluation.)
    """
    from optika_b200 import _util
    from oracle import grid as og

    rng = np.random.default_rng(5)
    ax = np.sort(rng.uniform(-0.3, 0.3, 1501))
    ay = np.sort(rng.uniform(-0.2, 0.4, 1203))
    field = na.Cartesian2dVectorArray(na.ScalarArray(ax, "fx"), na.ScalarArray(ay, "fy"))
    got = _util.direction(field).solid_angle_cell(("fx", "fy"))
    want = og.solid_angle_cell(*np.meshgrid(ax, ay, indexing="ij"))
    assert got.shape == {"fx": 1500, "fy": 1202}
    assert np.allclose(got.numpy(("fx", "fy")), want, rtol=1e-12, atol=2e-15)
    swapped = _util.direction(field).solid_angle_cell(("fy", "fx"))  # orientation flips the sign
    assert np.allclose(swapped.numpy(("fx", "fy")), -want, rtol=1e-12, atol=2e-15)
    # a leading (wavelength) axis: one grid per row
    rows = np.stack([ax[:40] * s for s in (1.0, 1.1, 1.3)])
    chromatic = na.Cartesian2dVectorArray(na.ScalarArray(rows, ("w", "fx")), na.ScalarArray(ay[:33], "fy"))
    got = _util.direction(chromatic).solid_angle_cell(("fx", "fy")).numpy(("w", "fx", "fy"))
    for k in range(3):
        assert np.allclose(got[k], og.solid_angle_cell(*np.meshgrid(rows[k], ay[:33], indexing="ij")), rtol=1e-12, atol=2e-15)
    # directions that are not normalised give the same solid angles
    d = _util.direction(field)
    scaled = na.Cartesian3dVectorArray(d.x * 3.0, d.y * 3.0, d.z * 3.0)
    assert np.allclose(scaled.solid_angle_cell(("fx", "fy")).numpy(("fx", "fy")), want, rtol=1e-12, atol=2e-15)


def test_detector_noise_is_seeded_per_block_of_pixels():
    """
    Shot and read noise (``ImagingSensor.expose``) come from one NumPy generator per block of 2^20 pixels, spawned from
    the seed: reproducible, independent of how many threads drew them, and with the right moments.
    """
    from optika_b200 import sensors

    lam = np.random.default_rng(0).uniform(0, 300, 3 * sensors._NOISE_BLOCK + 12345)
    a, b = sensors.shot_noise(lam, 11), sensors.shot_noise(lam, 11)
    assert a.dtype == np.int64 and np.array_equal(a, b) and not np.array_equal(a, sensors.shot_noise(lam, 12))
    # the first block does not depend on what follows it
    assert np.array_equal(sensors.shot_noise(lam[: sensors._NOISE_BLOCK], 11), a[: sensors._NOISE_BLOCK])
    z = (a - lam)[lam > 30] / np.sqrt(lam[lam > 30])
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert sensors.shot_noise(np.array([-3.0, 0.0]), 1).tolist() == [0, 0]
    r = sensors.read_noise(a, 2.5, 11) - a
    assert abs(r.mean()) < 1e-2 and abs(r.std() - 2.5) < 1e-2
    assert np.array_equal(sensors.read_noise(a.reshape(3, -1)[:, :7], 0.0, 4), a.reshape(3, -1)[:, :7])
    # blocks are independent streams: no block repeats another
    assert not np.array_equal((sensors.read_noise(np.zeros(2 * sensors._NOISE_BLOCK), 1.0, 3)[: sensors._NOISE_BLOCK]),
                              (sensors.read_noise(np.zeros(2 * sensors._NOISE_BLOCK), 1.0, 3)[sensors._NOISE_BLOCK:]))


def test_solid_angles_add_up():
    """Properties of the cell solid angles: a grid over the whole sphere sums to 4 pi, and refining a grid conserves the total."""
    from optika_b200 import _util

    def total(ax, ay):
        f = na.Cartesian2dVectorArray(na.ScalarArray(ax, "a"), na.ScalarArray(ay, "b"))
        return np.abs(_util.direction(f).solid_angle_cell(("a", "b")).numpy(("a", "b"))).sum()

    assert np.isclose(total(np.linspace(-np.pi, np.pi, 721), np.linspace(-np.pi / 2, np.pi / 2, 361)), 4 * np.pi, rtol=1e-12)
    # great-circle quadrilaterals are not the curvilinear cells: the total over a patch changes at second order in the
    # cell size only, and converges to the patch's d(azimuth) d(sin elevation)
    coarse = total(np.linspace(0.1, 0.3, 11), np.linspace(-0.2, 0.25, 12))
    fine = total(np.linspace(0.1, 0.3, 401), np.linspace(-0.2, 0.25, 441))
    exact = 0.2 * (np.sin(0.25) + np.sin(0.2))
    assert abs(fine - exact) < 1e-8 and abs(coarse - exact) < 2e-5 and abs(fine - exact) < abs(coarse - exact)
