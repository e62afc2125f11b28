"""
The Monte-Carlo electron kernel on the device (``optk_electrons_measured``) against the
oracle restatement of ``_electrons_measured_numba`` (``oracle/detector.py``) COUNT FOR COUNT -- both draw from the
same counter-based stream -- and against the reference's own pins
(``optika/sensors/materials/_ramanathan_2020/_ramanathan_2020_test.py:100-262``).
"""

import numpy as np
import pytest

from optika_b200 import named as na, sensors, units as u
from oracle import detector as od

pytestmark = pytest.mark.gpu

AXIS_XY = ("pixel_x", "pixel_y")


def oracle_planes(wavelength, temperature=300.0, **kwargs):
    """The per-plane records the product builds, for the oracle."""
    w = np.atleast_1d(np.asarray(wavelength, dtype=float))
    n, p = sensors.probability_of_n_pairs(na.ScalarArray(w, "_plane"), temperature)
    fano_inf = float(sensors.fano_factor(1.2398419843320026e-3 / 1000.0, temperature).ndarray)
    out = []
    for i in range(len(w)):
        out.append(dict(
            energy=1.2398419843320026e-3 / w[i], p_n=p[i], n=n, energy_pair_inf=float(sensors.energy_pair_inf(temperature)),
            fano_inf=fano_inf, **{k: (v[i] if np.ndim(v) else v) for k, v in kwargs.items()},
        ))
    return out


@pytest.mark.parametrize("wrap", [False, True])
@pytest.mark.parametrize(
    "wavelength,cce,implant",
    [
        (500 * u.nm, 1.0, 0.0),          # one pair per photon, no thinning
        (30 * u.nm, 0.5, 2000 * u.AA),   # 41 eV: pair-number table, collection efficiency ramp
        (0.21 * u.nm, 0.7, 2000 * u.AA), # 5.9 keV: rounded normal with Fano variance, ~1600 electrons per photon
    ],
)
def test_device_equals_the_oracle_count_for_count(cuda_device, wavelength, cce, implant, wrap):
    rng = np.random.default_rng(11)
    n_x, n_y = 7, 5
    many = wavelength > 1 * u.nm
    photons = rng.poisson(6 if many else 2, size=(n_x, n_y)).astype(np.int64)
    photons[3, 2] = 40 if many else 5
    absorption = 1 / u.um
    kwargs = dict(
        absorption=absorption, thickness_implant=implant, thickness_depletion=4 * u.um, thickness_substrate=14 * u.um,
        width_pixel_x=4 * u.um, width_pixel_y=6 * u.um, cce_backsurface=cce,
    )
    got = sensors.electrons_measured(
        na.ScalarArray(photons, AXIS_XY), wavelength, absorption=absorption, thickness_implant=implant,
        thickness_depletion=4 * u.um, thickness_substrate=14 * u.um,
        width_pixel=na.Cartesian2dVectorArray(4 * u.um, 6 * u.um), cce_backsurface=cce, axis_xy=AXIS_XY, wrap=wrap, seed=5,
    )
    (plane,) = oracle_planes(wavelength, **kwargs)
    want = od.electrons_measured(photons, plane, wrap=wrap, seed=5)
    assert got.axes == AXIS_XY
    assert want.sum() > 0
    mismatch = int((got.ndarray != want).sum())
    # identical streams and formulas; a libm difference in log / sincos can move an electron across a pixel
    # boundary once in ~1e13 draws
    assert mismatch == 0, f"{mismatch} pixels differ; totals {got.ndarray.sum()} vs {want.sum()}"


def test_planes_and_realisations(cuda_device):
    """Wavelength along its own axis (one record per image plane), `shape_random` adds independent realisations."""
    rng = np.random.default_rng(2)
    photons = na.ScalarArray(rng.poisson(3, size=(4, 6)).astype(np.int64), AXIS_XY)
    wavelength = na.ScalarArray(np.array([500 * u.nm, 30 * u.nm]), "wavelength")
    kwargs = dict(thickness_implant=2000 * u.AA, thickness_depletion=3 * u.um, thickness_substrate=14 * u.um,
                  width_pixel=15 * u.um, cce_backsurface=0.5)
    got = sensors.electrons_measured(photons, wavelength, axis_xy=AXIS_XY, shape_random=dict(experiment=3), seed=9, **kwargs)
    assert got.shape == {"wavelength": 2, "experiment": 3, "pixel_x": 4, "pixel_y": 6}
    assert np.all(got.ndarray >= 0)  # _ramanathan_2020_test.py:162
    k = np.imag(__import__("optika_b200").chemicals.Chemical("Si").n(wavelength).ndarray)
    absorption = 4 * np.pi * k / wavelength.ndarray
    planes = oracle_planes(
        np.repeat(wavelength.ndarray, 3), absorption=np.repeat(absorption, 3), thickness_implant=2000 * u.AA,
        thickness_depletion=3 * u.um, thickness_substrate=14 * u.um, width_pixel_x=15 * u.um, width_pixel_y=15 * u.um,
        cce_backsurface=0.5,
    )
    for i, plane in enumerate(planes):
        want = od.electrons_measured(photons.ndarray, plane, wrap=False, seed=9, plane_index=i)
        assert np.array_equal(got.ndarray.reshape(6, 4, 6)[i], want), i
    realisations = got.ndarray[1]
    assert not np.array_equal(realisations[0], realisations[1])  # independent draws
    # without pixel axes every element is its own 1 x 1 sensor (:632-636): nothing can leave it when wrapping
    alone = sensors.electrons_measured(photons, 500 * u.nm, wrap=True, seed=1, **kwargs)
    assert alone.shape == {"pixel_x": 4, "pixel_y": 6}
    assert np.all(alone.ndarray <= photons.ndarray) and alone.ndarray.sum() > 0.4 * photons.ndarray.sum()


def test_reference_pins_on_the_device(cuda_device):
    # _ramanathan_2020_test.py:180-232: the spread of the diffused charge is the analytic charge-diffusion width
    num = 41
    photons = np.zeros((num, num), dtype=np.int64)
    photons[num // 2, num // 2] = 20000
    electrons = sensors.electrons_measured(
        na.ScalarArray(photons, AXIS_XY), 500 * u.nm, absorption=1 / u.um, thickness_implant=0.0, thickness_depletion=0.0,
        thickness_substrate=14 * u.um, width_pixel=3 * u.um, cce_backsurface=1, axis_xy=AXIS_XY,
    ).ndarray
    assert electrons[num // 2, num // 2] < electrons.sum()
    offset = (np.arange(num) - num // 2) * 3 * u.um
    total = electrons.sum()
    mean_x, mean_y = (electrons * offset[:, None]).sum() / total, (electrons * offset[None, :]).sum() / total
    var = ((electrons * np.square(offset[:, None] - mean_x)).sum() + (electrons * np.square(offset[None, :] - mean_y)).sum()) / total
    expected = float(sensors.charge_diffusion(1 / u.um, 14 * u.um, 0.0))
    assert np.allclose(np.sqrt(var / 2), expected, rtol=0.05)
    # :235-262: a wrapped 3 x 3 grid keeps all the charge, a dropping one loses some
    small = np.zeros((3, 3), dtype=np.int64)
    small[1, 1] = 5000
    common = dict(absorption=1 / u.um, thickness_implant=0.0, thickness_depletion=0.0, thickness_substrate=14 * u.um,
                  width_pixel=2 * u.um, cce_backsurface=1, axis_xy=AXIS_XY)
    drop = sensors.electrons_measured(na.ScalarArray(small, AXIS_XY), 500 * u.nm, wrap=False, **common).ndarray.sum()
    wrapped = sensors.electrons_measured(na.ScalarArray(small, AXIS_XY), 500 * u.nm, wrap=True, **common).ndarray.sum()
    assert wrapped > drop


def test_parameters_along_the_pixel_axes_are_refused(cuda_device):
    photons = na.ScalarArray(np.ones((2, 2), dtype=np.int64), AXIS_XY)
    with pytest.raises(NotImplementedError, match="one value per image plane"):
        sensors.electrons_measured(photons, 500 * u.nm, cce_backsurface=na.ScalarArray(np.ones((2, 2)), AXIS_XY), axis_xy=AXIS_XY)
