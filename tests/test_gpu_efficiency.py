"""
Parity of the non-unit efficiencies evaluated inside the trace kernel (measured mirrors and
rulings: ``numpy.interp`` tables; groove profiles of ``optika/rulings/_rulings.py:404-1073``)
against the oracle, as unit operations and inside a full system trace.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import transformations as tf
from oracle import raytrace as ora

import configs
import parity
from test_oracle_efficiency import measured

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(11)


def random_rays(n=5000, wavelength=(150 * u.AA, 900 * u.AA)):
    d = rng.normal(size=(3, n)) * 0.2
    d[2] = 1.0
    d /= np.linalg.norm(d, axis=0)
    normal = rng.normal(size=(3, n)) * 0.1
    normal[2] = -1.0
    normal /= np.linalg.norm(normal, axis=0)
    ax = "ray"
    rays = optika.rays.RayVectorArray(
        wavelength=na.ScalarArray(rng.uniform(*wavelength, n), ax),
        position=na.Cartesian3dVectorArray(
            na.ScalarArray(rng.uniform(-8, 8, n), ax), na.ScalarArray(rng.uniform(-8, 8, n), ax), 0.0
        ),
        direction=na.Cartesian3dVectorArray(*[na.ScalarArray(c, ax) for c in d]),
    )
    nrm = na.Cartesian3dVectorArray(*[na.ScalarArray(c, ax) for c in normal])
    return rays, nrm, normal


SPACINGS = {
    "constant": 1 * u.um,
    "polynomial": optika.rulings.Polynomial1dRulingSpacing(
        coefficients={0: 1 * u.um, 1: 2e-5, 2: 1e-6 / u.mm}, normal=na.Cartesian3dVectorArray(1, 0, 0)
    ),
}


@pytest.mark.parametrize("spacing", list(SPACINGS))
@pytest.mark.parametrize("order", [-2, -1, 0, 1, 2, 3])
@pytest.mark.parametrize("kind", ["Sinusoidal", "Square", "Sawtooth", "Triangular", "Rectangular"])
def test_groove_profile_efficiency(cuda_device, kind, order, spacing):
    cls = getattr(optika.rulings, kind + "Rulings")
    kwargs = dict(spacing=SPACINGS[spacing], depth=18 * u.nm, diffraction_order=order)
    if kind == "Rectangular":
        kwargs["ratio_duty"] = 0.35
    rulings = cls(**kwargs)
    rays, normal, n_nd = random_rays()
    got = rulings.efficiency(rays, normal)
    r0, _ = configs.flatten_rays(rays)
    want = ora.rulings_efficiency(rulings, r0, tuple(n_nd))
    assert got.shape == {"ray": 5000}
    assert np.allclose(got.ndarray, want, rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("order", [0, 1, 4, 9, -3])
def test_bessel_function_over_a_wide_argument_range(cuda_device, order):
    import scipy.special

    n = 4000
    depth = 1 * u.um
    x = np.concatenate([np.geomspace(1e-12, 1, n // 4), np.linspace(1, 400, 3 * n // 4)])  # 2 gamma
    w = 2 * np.pi * depth / x  # gamma = pi depth / (w cos), cos = 1
    rays = optika.rays.RayVectorArray(
        wavelength=na.ScalarArray(w, "ray"), direction=na.Cartesian3dVectorArray(0.0, 0.0, 1.0)
    )
    rulings = optika.rulings.SinusoidalRulings(spacing=10 * u.um, depth=depth, diffraction_order=order)
    got = rulings.efficiency(rays, na.Cartesian3dVectorArray(0.0, 0.0, -1.0)).ndarray
    want = scipy.special.jv(order, 2 * (np.pi * depth / w))
    assert np.abs(got - want).max() < 5e-14


def test_measured_efficiencies(cuda_device):
    w = np.linspace(200 * u.AA, 800 * u.AA, 13)
    table = measured(rng.uniform(0.05, 0.9, 13), w)
    mirror = optika.materials.MeasuredMirror(table)
    rulings = optika.rulings.MeasuredRulings(
        spacing=1 * u.um, diffraction_order=1, efficiency_measured=measured(rng.uniform(0.1, 0.5, 13)[::-1], w[::-1])
    )
    rays, normal, n_nd = random_rays(wavelength=(150 * u.AA, 900 * u.AA))  # beyond both ends: clamped
    rays.wavelength.ndarray[:13] = w  # exactly on the knots
    r0, _ = configs.flatten_rays(rays)
    got_m = mirror.efficiency(rays, normal).ndarray
    got_r = rulings.efficiency(rays, normal).ndarray
    assert np.allclose(got_m, ora.material_efficiency(mirror, r0, None), rtol=1e-12, atol=0)
    assert np.allclose(got_r, ora.rulings_efficiency(rulings, r0, None), rtol=1e-12, atol=0)
    assert np.array_equal(got_m[:13], table.outputs.ndarray)


def grating_system(rulings, mirror, num_pixel=256):
    """cfg 2 geometry with a lossy grating."""
    system = configs.spherical_grating(num_field=5, num_pupil=12, num_wavelength=5, num_pixel=num_pixel)
    grating = system.surfaces[0]
    grating.rulings = rulings
    grating.material = mirror
    return system


def test_system_trace_with_efficiencies(cuda_device):
    w = np.linspace(150 * u.AA, 650 * u.AA, 21)
    mirror = optika.materials.MeasuredMirror(measured(np.exp(-np.square((w - 400 * u.AA) / (150 * u.AA))), w))
    rulings = optika.rulings.SawtoothRulings(spacing=(1 / 1200) * u.mm, depth=12 * u.nm, diffraction_order=1)
    system = grating_system(rulings, mirror)
    result = system.raytrace(accumulate=True, **configs.PHYSICAL)
    _, rays0 = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays0)
    states = ora.accumulate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
    out = result.outputs
    order = tuple(rays0.shape)  # the oracle's ray order
    get = lambda a: na.as_named_array(a).numpy(("surface",) + order).reshape(len(system.surfaces_all), -1)  # noqa: E731
    got = dict(
        wavelength=get(out.wavelength), px=get(out.position.x), py=get(out.position.y), pz=get(out.position.z),
        dx=get(out.direction.x), dy=get(out.direction.y), dz=get(out.direction.z), intensity=get(out.intensity),
        attenuation=get(out.attenuation), index_refraction=get(out.index_refraction),
        unvignetted=get(out.unvignetted).astype(bool),
    )
    parity.compare_states(got, states, system.surfaces_all)
    assert 0 < states["intensity"][-1].max() < 1 and np.ptp(states["intensity"][-1]) > 0.01


def test_fused_image_with_efficiencies_and_a_configuration_axis(cuda_device):
    from oracle import binning as orb

    w = np.linspace(150 * u.AA, 650 * u.AA, 9)
    table = measured(np.zeros(9), w)
    table.outputs = na.ScalarArray(
        np.stack([np.linspace(0.2, 0.8, 9), np.linspace(0.95, 0.4, 9)]), ("coating", "wavelength_measured")
    )
    mirror = optika.materials.MeasuredMirror(table)
    assert mirror.shape == {"coating": 2}
    rulings = optika.rulings.SquareRulings(spacing=(1 / 1200) * u.mm, depth=9 * u.nm, diffraction_order=1)
    system = grating_system(rulings, mirror)
    edges = na.ScalarArray(np.array([100 * u.AA, 400 * u.AA, 700 * u.AA]), "wavelength")
    image = system.image_rays(edges, counts=True, **configs.PHYSICAL)
    flux = image.flux.cpu().numpy()
    assert flux.shape == (2, 2, 256, 256)
    _, rays0 = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays0)
    ex, ey = system.sensor.pixel_edges()
    for c in range(2):
        surfaces = ora.select_config(system.surfaces_all, {"coating": c})
        out = ora.propagate_rays(surfaces, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
        local = ora._rays_transform(surfaces[-1].transformation, out, inverse=True)
        want, _, _ = orb.collect(local, edges.ndarray, ex, ey)
        assert np.isclose(flux[c].sum(), want.sum(), rtol=1e-9) and want.sum() > 0
        assert (~np.isclose(flux[c], want, rtol=1e-9, atol=1e-9 * want.max())).sum() <= 8
    assert not np.isclose(flux[0].sum(), flux[1].sum(), rtol=1e-3)
