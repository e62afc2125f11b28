"""
Host-side models fitted to traced rays: the pins the reference's tests hold for them
-- a linear scene -> sensor map is reproduced and inverted to 1e-9 deg
(``optika/distortion/_distortion_test.py:40-44, 108-123``), ``inverse == 1 / model``
(``optika/radiometry/_vignetting_test.py:37-41``) -- and the least-squares machinery underneath.
"""

import numpy as np
import pytest

from optika_b200 import named as na, units as u
from optika_b200._polynomial import PolynomialFit, _exponents
from optika_b200.distortion import PolynomialDistortionModel
from optika_b200.radiometry import PolynomialVignettingModel, InterpolatedEffectiveAreaModel
from optika_b200.vectors import SpectralPositionalVectorArray


def _scene(num=5):
    return SpectralPositionalVectorArray(
        wavelength=na.linspace(500, 600, axis="wavelength", num=3) * u.nm,
        position=na.Cartesian2dVectorLinearSpace(
            start=-1 * u.deg, stop=+1 * u.deg, axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=num,
        ),
    )


@pytest.mark.parametrize("degree", [1, 2])
def test_distortion_roundtrip_of_a_linear_map(degree):
    scene = _scene()
    model = PolynomialDistortionModel(
        coordinates_scene=scene,
        coordinates_sensor=na.Cartesian2dVectorArray(x=scene.position.x * (10 / u.deg), y=scene.position.y * (10 / u.deg)),
        axis_wavelength="wavelength", axis_field=("field_x", "field_y"), degree=degree,
    )
    distorted = model.distort(scene)
    assert np.array_equal(distorted.wavelength.ndarray, scene.wavelength.ndarray)  # carried through unchanged
    result = model.undistort(distorted)
    ex = na.as_named_array(result.position.x - scene.position.x).ndarray
    ey = na.as_named_array(result.position.y - scene.position.y).ndarray
    assert np.all(np.sqrt(ex**2 + ey**2) < 1e-9 * u.deg)
    assert model.fit.coefficient_names is not None and model.fit_inverse.coefficient_names is not None
    assert np.nanmax(model.residual.ndarray) < 1e-12


def test_distortion_fit_recovers_a_quadratic_chromatic_map_and_respects_the_mask():
    scene = _scene(num=9)
    w = (na.as_named_array(scene.wavelength) - 550 * u.nm) / (50 * u.nm)
    x, y = scene.position.x / u.deg, scene.position.y / u.deg
    sensor = na.Cartesian2dVectorArray(x=10 * x + 0.3 * x * x + 0.05 * w * y + 0.2 * w, y=10 * y - 0.1 * x * y + 0.4 * w * w)
    where = na.ScalarArray(np.ones((3, 9, 9), dtype=bool), ("wavelength", "field_x", "field_y"))
    where.ndarray[1, 4, :] = False
    corrupted = na.Cartesian2dVectorArray(
        x=na.ScalarArray(np.where(where.ndarray, na.broadcast_to(sensor.x, where.shape).ndarray, 1e6), where.axes),
        y=na.broadcast_to(sensor.y, where.shape),
    )
    model = PolynomialDistortionModel(scene, corrupted, "wavelength", ("field_x", "field_y"), degree=2, where=where)
    assert np.nanmax(model.residual.ndarray) < 1e-10  # the masked (corrupted) points did not enter the fit
    assert np.isnan(model.residual.ndarray[1, 4]).all()
    linear = PolynomialDistortionModel(scene, sensor, "wavelength", ("field_x", "field_y"), degree=1)
    assert np.nanmax(linear.residual.ndarray) > 0.05  # deliberately underfit: a visible residual


def test_vignetting_model_and_its_inverse():
    scene = _scene(num=13)
    r2 = (scene.position.x / u.deg) ** 2 + (scene.position.y / u.deg) ** 2
    illumination = 1 - 0.1 * r2
    model = PolynomialVignettingModel(scene, illumination, "wavelength", ("field_x", "field_y"), degree=2)
    assert np.nanmax(np.abs(model.residual.ndarray)) < 1e-12
    value = model(scene)
    assert set(value.shape) == {"wavelength", "field_x", "field_y"}
    assert np.array_equal(model.inverse(scene).ndarray, (1 / value).ndarray)
    underfit = PolynomialVignettingModel(scene, illumination, "wavelength", ("field_x", "field_y"), degree=1)
    assert np.nanmax(np.abs(underfit.residual.ndarray)) > 0.05


def test_effective_area_interpolates_linearly_with_clamped_ends():
    wavelength = na.linspace(100, 1000, axis="wavelength", num=10)
    area = 10 * np.exp(-(((wavelength - 500) / 150) ** 2))
    model = InterpolatedEffectiveAreaModel(wavelength=wavelength, area=area, axis_wavelength="wavelength")
    fine = na.linspace(0, 1100, axis="wavelength", num=45)
    assert np.allclose(model(fine).ndarray, np.interp(fine.ndarray, wavelength.ndarray, area.ndarray), rtol=0, atol=0)
    # a configuration axis on the calibration areas is interpolated index by index
    both = na.ScalarArray(np.stack([area.ndarray, 2 * area.ndarray]), ("config", "wavelength"))
    two = InterpolatedEffectiveAreaModel(wavelength=wavelength, area=both, axis_wavelength="wavelength")(fine)
    assert two.shape == {"config": 2, "wavelength": 45}
    assert np.allclose(two.ndarray[1], 2 * two.ndarray[0])


def test_polynomial_fit_batches_over_axes_that_are_not_scene_axes():
    assert _exponents(2, 2) == [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]
    x = na.linspace(-1, 1, axis="x", num=7)
    c = na.ScalarArray(np.array([1.0, 2.0, 3.0]), "config")
    y = c * x * x + 0.5 * x - c
    fit = PolynomialFit(inputs=(x,), outputs=(y,), degree=2, axes=("x",))
    (prediction,) = fit.predictions
    assert prediction.shape == {"config": 3, "x": 7}
    assert np.allclose(prediction.ndarray, na.broadcast_to(y, prediction.shape).ndarray, atol=1e-13)
    (at,) = fit(na.ScalarArray(np.array([0.25]), "x"))
    assert np.allclose(at.ndarray[:, 0], np.array([1, 2, 3]) * 0.0625 + 0.125 - np.array([1, 2, 3]))
