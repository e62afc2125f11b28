"""
``SequentialSystem.distortion`` / ``vignetting`` / ``area_effective`` (``optika/systems/_sequential.py:1208-1512``,
SURVEY.md section 8f-4) on the device: the per-field-point reductions come out of the trace kernel, the fits are host
work.  Checked against the same quantities formed with NumPy from rays the ORACLE traced, and against the
geometry of the systems.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na, units as u
from oracle import raytrace as ora

import configs

pytestmark = pytest.mark.gpu


def chromatic_newtonian(num_field=5, num_pupil=16):
    system = configs.newtonian(num_field=num_field, num_pupil=num_pupil)
    system.grid_input.wavelength = na.linspace(450 * u.nm, 650 * u.nm, axis="wavelength", num=3)
    return system


def oracle_moments(system):
    """where / mean position / unvignetted fraction / summed intensity per (wavelength, field point), NumPy on oracle rays."""
    result, rays = system._input(None, None, None, None, False, False)
    r0, shape_ = configs.flatten_rays(rays)
    dims = tuple(shape_.values())
    out = ora.propagate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, converge=True, extended=True)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    names = list(shape_)
    k = tuple(names.index(ax) for ax in ("pupil_x", "pupil_y"))
    unv = local["unvignetted"].reshape(dims)
    where = unv.any(axis=k)
    use = unv | ~np.expand_dims(where, k)
    x = np.mean(local["px"].reshape(dims), axis=k, where=use)
    y = np.mean(local["py"].reshape(dims), axis=k, where=use)
    outer = tuple(ax for ax in names if ax not in ("pupil_x", "pupil_y"))
    return outer, where, x, y, unv.mean(axis=k), np.sum(local["intensity"].reshape(dims), axis=k, where=unv)


def aligned(a, axes):
    a = na.as_named_array(a)
    return np.transpose(a.ndarray, [a.axes.index(ax) for ax in axes])


def test_distortion_model_of_the_newtonian(cuda_device):
    system = chromatic_newtonian()
    model = system.distortion(**configs.PHYSICAL, degree=2)
    outer, where, x, y, _, _ = oracle_moments(system)
    assert np.array_equal(aligned(model.where, outer), where) and where.all()
    assert np.allclose(aligned(model.coordinates_sensor.x, outer), x, rtol=0, atol=1e-9 * np.abs(x).max())
    assert np.allclose(aligned(model.coordinates_sensor.y, outer), y, rtol=0, atol=1e-9 * np.abs(x).max())
    assert model.axis_wavelength == "wavelength" and set(model.axis_field) == {"field_x", "field_y"}
    # a 200 mm mirror images 0.1 deg to 0.35 mm; a quadratic model leaves only coma-sized residuals
    assert np.nanmax(model.residual.ndarray) < 2e-4
    scene = model.coordinates_scene
    back = model.undistort(model.distort(scene))
    error = np.hypot(na.as_named_array(back.position.x - scene.position.x).ndarray,
                     na.as_named_array(back.position.y - scene.position.y).ndarray)
    assert error.max() < 1e-3 * (0.1 * u.deg)
    # plate scale: d(sensor) / d(field) = focal length, up to the mirror flips of the fold
    names = model.fit.coefficient_names
    scale = model.fit._scale[0]
    linear = [abs(model.fit.coefficients[0, 0, names.index(f"x{j}^1")]) / scale[j] for j in (1, 2)]
    assert max(linear) == pytest.approx(200.0, rel=0.03)  # 203.6: f tan(theta) plus the third-order terms the fit absorbs


def test_vignetting_model_of_the_toroidal_spectrograph(cuda_device):
    system = configs.toroidal_vls(num_field=7, num_pupil=24, num_wavelength=3)  # the octagon clips ~40 % of the rays
    model = system.vignetting(**configs.PHYSICAL, degree=2)
    outer, where, _, _, illumination, _ = oracle_moments(system)
    k = tuple(outer.index(ax) for ax in ("field_x", "field_y"))
    want = illumination / np.mean(illumination, axis=k, where=where, keepdims=True)
    assert np.array_equal(aligned(model.where, outer), where)
    assert np.allclose(aligned(model.illumination, outer), want, rtol=1e-12, atol=0)
    assert 0.2 < illumination.mean() < 0.9
    value = model(model.coordinates_scene)
    assert np.array_equal(model.inverse(model.coordinates_scene).ndarray, (1 / value).ndarray)


def test_vignetting_excludes_field_points_without_a_surviving_ray(cuda_device):
    system = chromatic_newtonian(num_field=5, num_pupil=12)
    field = na.Cartesian2dVectorLinearSpace(
        -2.0 * u.deg, 2.0 * u.deg, axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=5, centers=True)
    model = system.vignetting(field=field, **configs.PHYSICAL, degree=1)
    where = model.where.ndarray
    assert where.any() and not where.all()
    masked = np.where(where, model.illumination.ndarray, np.nan)
    axes = model.illumination.axes
    k = tuple(axes.index(ax) for ax in ("field_x", "field_y"))
    assert np.allclose(np.nanmean(masked, axis=k), 1.0)  # unit mean over the field points that enter (:1355-1358)


def test_models_need_a_wavelength_axis(cuda_device):
    system = configs.newtonian(num_field=3, num_pupil=4)  # scalar wavelength
    with pytest.raises(ValueError, match="wavelength grid to vary along its own logical axis"):
        system.distortion(**configs.PHYSICAL)
    with pytest.raises(ValueError, match="only one wavelength axis"):
        system.area_effective(normalized_field=False, normalized_pupil=False,
                              pupil=na.Cartesian2dVectorArray(na.linspace(-40, 40, axis="px", num=5),
                                                              na.linspace(-40, 40, axis="py", num=5)))


def test_effective_area_of_the_newtonian_is_its_clear_aperture(cuda_device):
    system = chromatic_newtonian(num_field=3, num_pupil=4)
    pupil = na.Cartesian2dVectorArray(
        x=na.linspace(-44 * u.mm, 44 * u.mm, axis="_pupil_x", num=89), y=na.linspace(-44 * u.mm, 44 * u.mm, axis="_pupil_y", num=89)
    )
    model = system.area_effective(pupil=pupil, normalized_field=False, normalized_pupil=False, seed=1)
    assert model.area.shape == {"wavelength": 3}
    # 80 x 80 mm mirror minus the shadow of the 50 x 50 mm obscuration, which is tilted by 45 deg about y
    # (50 cos 45 deg x 50 mm), perfect mirrors: 6400 - 2500 / sqrt(2) = 4632 mm^2, to the sampling error of
    # 88 x 88 one-millimetre cells along the edges
    clear = 6400.0 - 2500.0 / np.sqrt(2.0)
    assert np.allclose(model.area.ndarray, clear, rtol=0.01)
    # another seed moves the stratified samples, not the answer
    other = system.area_effective(pupil=pupil, normalized_field=False, normalized_pupil=False, seed=2)
    assert np.allclose(other.area.ndarray, model.area.ndarray, rtol=0.02)
    assert not np.array_equal(other.area.ndarray, model.area.ndarray)
    assert np.allclose(model(na.ScalarArray(np.array([500e-6]), "wavelength")).ndarray, model.area.ndarray[:1], rtol=0.02)


def test_effective_area_in_normalised_pupil_coordinates(cuda_device):
    system = chromatic_newtonian(num_field=3, num_pupil=4)
    model = system.area_effective()  # reference defaults: normalised field and pupil, 11 x 11 vertices over [-1, 1]^2
    # the normalised pupil spans the stop (the 80 x 80 mm primary): the cells cover it exactly and the
    # obscuration removes what falls on the tilted 50 x 50 mm fold; 10 x 10 cells of 8 mm resolve that to ~15 %
    clear = 6400.0 - 2500.0 / np.sqrt(2.0)
    assert np.all(model.area.ndarray > 0.8 * clear) and np.all(model.area.ndarray < 1.2 * clear)
