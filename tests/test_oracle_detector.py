"""
Detector physics after binning: the oracle (``oracle/detector.py``) against every pin
the reference's tests hold for this path, and the product's host functions against the oracle.

* ``optika/sensors/materials/_diffusion_test.py``: ``charge_diffusion > 0``, ``0 < mean_charge_capture < 1``,
  the kernel has the named axes;
* ``optika/sensors/materials/_ramanathan_2020/_ramanathan_2020_test.py:180-262``: the spread of the charge
  diffused from one pixel equals ``charge_diffusion`` within 5 %; a wrapped grid holds more charge than a
  dropping one; electron counts are non-negative.
"""

import numpy as np
import pytest

from optika_b200 import named as na, sensors, units as u
from oracle import detector as od


def plane(energy=2.48, absorption=1 / u.um, implant=0.0, depletion=0.0, substrate=14 * u.um, pixel=3 * u.um, cce=1.0,
          p_n=None):
    if p_n is None:
        p_n = np.zeros(20)
        p_n[0] = 1.0  # one pair per photon (visible light)
    return dict(
        energy=energy, absorption=absorption, thickness_implant=implant, thickness_depletion=depletion,
        thickness_substrate=substrate, width_pixel_x=pixel, width_pixel_y=pixel, cce_backsurface=cce,
        p_n=p_n, n=np.arange(1, 21), energy_pair_inf=3.65, fano_inf=0.12,
    )


def spread(electrons, width_pixel):
    num = electrons.shape[0]
    offset = (np.arange(num) - num // 2) * width_pixel
    total = electrons.sum()
    mean_x = (electrons * offset[:, None]).sum() / total
    mean_y = (electrons * offset[None, :]).sum() / total
    var_x = (electrons * np.square(offset[:, None] - mean_x)).sum() / total
    var_y = (electrons * np.square(offset[None, :] - mean_y)).sum() / total
    return np.sqrt((var_x + var_y) / 2)


def test_oracle_diffusion_spread_matches_the_analytic_width():
    # _ramanathan_2020_test.py:180-232, same numbers
    num = 41
    photons = np.zeros((num, num), dtype=np.int64)
    photons[num // 2, num // 2] = 20000
    electrons = od.electrons_measured(photons, plane(), wrap=False, seed=7)
    assert electrons[num // 2, num // 2] < electrons.sum()
    expected = od.charge_diffusion(1 / u.um, 14 * u.um, 0.0)
    assert np.allclose(spread(electrons, 3 * u.um), expected, rtol=0.05)
    assert (electrons >= 0).all()


def test_oracle_wrapped_grid_keeps_more_charge():
    # _ramanathan_2020_test.py:235-262
    photons = np.zeros((3, 3), dtype=np.int64)
    photons[1, 1] = 5000
    p = plane(pixel=2 * u.um)
    drop = od.electrons_measured(photons, p, wrap=False, seed=1).sum()
    wrapped = od.electrons_measured(photons, p, wrap=True, seed=1).sum()
    assert wrapped > drop and wrapped == 5000


def test_oracle_counts_are_exact_where_nothing_is_random():
    # no field-free region (fully depleted), unit collection efficiency, a delta pair-number distribution
    p_n = np.zeros(20)
    p_n[2] = 1.0  # three pairs per photon
    photons = np.arange(12, dtype=np.int64).reshape(3, 4)
    electrons = od.electrons_measured(photons, plane(depletion=14 * u.um, p_n=p_n), wrap=False, seed=3)
    assert np.array_equal(electrons, 3 * photons)
    # collection efficiency h0 at the back surface thins the pairs binomially
    thin = od.electrons_measured(np.full((1, 1), 20000), plane(depletion=14 * u.um, implant=1.0, cce=0.25, absorption=1e4), False, 5)
    assert abs(thin[0, 0] / 20000 - 0.25) < 0.02  # every photon is absorbed right at the surface: h = h0


@pytest.mark.parametrize("width_diffusion", [10 * u.um, na.linspace(1, 10, "width", 5) * u.um])
def test_host_closed_forms_match_the_oracle(width_diffusion):
    width_pixel = 15 * u.um
    mcc = sensors.mean_charge_capture(width_diffusion, width_pixel)
    assert np.all(mcc.ndarray > 0) and np.all(mcc.ndarray < 1)  # _diffusion_test.py:59-60
    w = na.as_named_array(width_diffusion).ndarray
    assert np.allclose(mcc.ndarray, od.mean_charge_capture(w, width_pixel), rtol=1e-14)
    kernel = sensors.kernel_diffusion(width_diffusion, width_pixel, "x", "y")
    assert kernel.outputs.shape["x"] == 3 and kernel.outputs.shape["y"] == 3  # _diffusion_test.py:86-88
    for k, wk in enumerate(np.atleast_1d(w)):
        got = kernel.outputs.ndarray if np.ndim(w) == 0 else kernel.outputs[{"width": k}].numpy(("x", "y"))
        got = np.asarray(got).reshape(3, 3) if np.ndim(w) == 0 else got
        assert np.allclose(got, od.kernel_diffusion(wk, width_pixel), rtol=1e-13)
        assert got.sum() < 1 and got[1, 1] == got.max()
    sigma = sensors.charge_diffusion(1 / u.um, 15 * u.um, 5 * u.um)
    assert sigma > 0 and np.isclose(sigma, od.charge_diffusion(1 / u.um, 15 * u.um, 5 * u.um), rtol=1e-15)


def test_pair_creation_model():
    # Ramanathan & Kurinsky 2020: E_g(300 K) = 1.123 eV, 3.65 eV per pair asymptotically, Fano 0.12
    assert np.isclose(sensors.energy_bandgap(300.0), 1.1230, atol=1e-3)
    assert np.isclose(sensors.energy_pair_inf(300.0), 3.646, atol=2e-3)
    assert np.isclose(sensors.fano_factor_inf(300.0), 0.1164, atol=2e-3)
    # a 5.9 keV photon (Fe-55) makes ~1600 electrons; visible light exactly one
    assert np.isclose(float(sensors.quantum_yield_ideal(1.2398419843320026e-3 / 5900.0).ndarray), 5900 / 3.646, rtol=2e-3)
    assert np.isclose(float(sensors.quantum_yield_ideal(500 * u.nm).ndarray), 1.0, atol=1e-3)  # tables interpolated between rows
    n, p = sensors.probability_of_n_pairs(na.ScalarArray(np.array([500 * u.nm, 100 * u.nm, 30 * u.nm]), "w"))
    assert p.shape == (3, 20) and np.allclose(p.sum(-1), 1.0, atol=1e-3)
    assert p[0, 0] == pytest.approx(1.0, abs=1e-3) and (p[2] * n).sum() > (p[1] * n).sum() > 1.5
    f = sensors.fano_factor(na.ScalarArray(np.array([100 * u.nm, 1 * u.nm]), "w")).ndarray
    assert 0 < f[0] < 0.5 and np.isclose(f[1], sensors.fano_factor_inf(300.0))
