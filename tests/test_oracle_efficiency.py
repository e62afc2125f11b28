"""
Oracle restatement of the non-unit efficiencies (``optika/materials/_materials.py:279-305``,
``optika/rulings/_rulings.py:287-313, 404-1073``) against what the reference's own tests pin
(``optika/rulings/_rulings_test.py:63-75``: ``0 <= efficiency <= 1``) and against closed-form
identities of the formulas, plus the host lowering of those elements.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import _lib, _lowering
from oracle import raytrace as ora

rng = np.random.default_rng(5)


def rays_and_normal(n=2000, wavelength=(200 * u.AA, 700 * u.AA)):
    d = rng.normal(size=(3, n)) * 0.15
    d[2] = 1.0
    d /= np.linalg.norm(d, axis=0)
    rays = ora.make_rays(n, wavelength=rng.uniform(*wavelength, n), dx=d[0], dy=d[1], dz=d[2],
                         px=rng.uniform(-5, 5, n), py=rng.uniform(-5, 5, n))
    normal = (np.zeros(n), np.zeros(n), -np.ones(n))
    return rays, normal


PROFILES = [
    optika.rulings.SquareRulings(spacing=1 * u.um, depth=15 * u.nm, diffraction_order=1),
    optika.rulings.SawtoothRulings(spacing=1 * u.um, depth=15 * u.nm, diffraction_order=1),
    optika.rulings.TriangularRulings(spacing=1 * u.um, depth=15 * u.nm, diffraction_order=1),
    optika.rulings.RectangularRulings(spacing=1 * u.um, depth=15 * u.nm, ratio_duty=0.3, diffraction_order=1),
    optika.rulings.RectangularRulings(spacing=1 * u.um, depth=15 * u.nm, ratio_duty=0.5, diffraction_order=0),
]


@pytest.mark.parametrize("rulings", PROFILES)
def test_efficiency_is_a_fraction(rulings):
    rays, normal = rays_and_normal()
    e = ora.rulings_efficiency(rulings, rays, normal)
    assert np.all(e >= 0) and np.all(e <= 1)  # _rulings_test.py:74-75


def test_ideal_rulings_have_unit_efficiency():
    rays, normal = rays_and_normal(10)
    assert ora.rulings_efficiency(optika.rulings.Rulings(spacing=1 * u.um, diffraction_order=2), rays, normal) == 1.0


def test_square_profile_orders_sum_to_one():
    rays, normal = rays_and_normal(50)
    total = 0.0
    for m in range(-2001, 2002):
        total = total + ora.rulings_efficiency(
            optika.rulings.SquareRulings(spacing=1 * u.um, depth=25 * u.nm, diffraction_order=m), rays, normal
        )
    assert np.allclose(total, 1.0, atol=5e-4)


def test_sinusoidal_profile_is_the_unsquared_bessel_function():
    import scipy.special

    n = 100
    rays = ora.make_rays(n, wavelength=np.linspace(100 * u.AA, 900 * u.AA, n), dz=np.ones(n))  # normal incidence
    normal = (np.zeros(n), np.zeros(n), -np.ones(n))
    r = optika.rulings.SinusoidalRulings(spacing=0.4 * u.um, depth=15 * u.nm, diffraction_order=1)
    e = ora.rulings_efficiency(r, rays, normal)
    assert np.allclose(e, scipy.special.jv(1, 2 * np.pi * 15 * u.nm / rays["wavelength"]), rtol=1e-13)
    assert e.min() < 0  # unsquared: the reference returns the amplitude (_rulings.py:455)


def measured(values, wavelengths, axis="wavelength_measured"):
    w = na.ScalarArray(np.asarray(wavelengths, dtype=float), axis)
    return na.FunctionArray(
        inputs=optika.vectors.SpectralDirectionalVectorArray(wavelength=w, direction=na.Cartesian3dVectorArray(0, 0, 1)),
        outputs=na.ScalarArray(np.asarray(values, dtype=float), axis),
    )


def test_measured_mirror_interpolates_and_clamps():
    center, width = 304 * u.AA, 10 * u.AA
    w = np.linspace(center - 3 * width, center + 3 * width, 11)
    table = measured(np.exp(-np.square((w - center) / width) / 2), w)  # the docstring example, _materials.py:196-219
    mirror = optika.materials.MeasuredMirror(table)
    assert mirror.is_mirror and mirror.shape == {}
    x = np.array([w[0] - 1e-6, w[0], 0.5 * (w[3] + w[4]), w[5], w[-1], w[-1] + 1e-6])
    e = ora.material_efficiency(mirror, ora.make_rays(len(x), wavelength=x), None)
    assert e[0] == e[1] == table.outputs.ndarray[0] and e[-1] == e[-2] == table.outputs.ndarray[-1]
    assert e[3] == 1.0
    assert np.isclose(e[2], 0.5 * (table.outputs.ndarray[3] + table.outputs.ndarray[4]))


def test_lowering_of_efficiency_elements():
    w = np.linspace(200 * u.AA, 400 * u.AA, 7)
    mirror = optika.materials.MeasuredMirror(measured(np.linspace(0.1, 0.7, 7)[::-1], w[::-1]))
    surface = optika.surfaces.Surface(
        material=mirror,
        rulings=optika.rulings.RectangularRulings(
            spacing=2 * u.um, depth=na.ScalarArray(np.array([10.0, 20.0]) * u.nm, "cfg"), ratio_duty=0.25,
            diffraction_order=-1,
        ),
    )
    table, shape_ = _lowering.lower_system([surface])
    assert shape_ == {"cfg": 2} and len(table) == 2 and len(table.keep) == 4
    for c, depth in enumerate((10e-6, 20e-6)):
        S = table[c]
        assert S.material_kind == _lib.MAT_MIRROR and S.material_efficiency == _lib.EFF_LUT and S.material_lut_n == 7
        assert S.ruling_profile == _lib.PROFILE_RECTANGULAR and S.ruling_order == -1 and S.ruling_duty == 0.25
        assert np.isclose(S.ruling_depth, depth)
    x = np.ctypeslib.as_array((_lib.C.c_double * 7).from_address(table[0].material_lut_x))
    y = np.ctypeslib.as_array((_lib.C.c_double * 7).from_address(table[0].material_lut_y))
    assert np.array_equal(x, w) and np.allclose(y, np.linspace(0.1, 0.7, 7))  # sorted ascending for numpy.interp
    with pytest.raises(ValueError):
        bad = measured(np.ones((2, 3)).ravel(), np.arange(6.0))
        bad.inputs.direction = na.Cartesian3dVectorArray(na.ScalarArray(np.zeros(2), "angle"), 0, 1)
        _lowering.lower_system([optika.surfaces.Surface(material=optika.materials.MeasuredMirror(bad))])


def test_measured_tables_need_a_device():
    """No CPU fallback: the tables live in device memory, so creating the system needs CUDA."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from optika_b200 import _engine

    mirror = optika.materials.MeasuredMirror(measured([0.5, 0.6], [1e-5, 2e-5]))
    with pytest.raises(_lib.OptkError):
        _engine.CompiledSystem([optika.surfaces.Surface(material=mirror)])
