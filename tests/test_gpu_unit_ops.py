"""
The reference's unit operations, executed on the device by one-surface traces with a
restricted stage mask (SURVEY.md section 8b "secondary seams"), against the oracle:
``sag.__call__ / normal / intercept / propagate_rays`` (``optika/sags/_abc.py:48-122``),
``aperture.__call__ / clip_rays`` (``optika/apertures/_apertures.py:69-102``),
``spacing.__call__`` (``optika/rulings/_spacing.py:27-41``),
``rulings.incident_effective`` (``optika/rulings/_rulings.py:170-204``) and
``snells_law`` (``optika/materials/_snells_law.py:41-47``).
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import transformations as tf
from oracle import raytrace as ora

import configs

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(3)
N = 2000
AX = "ray"


def points(spread=20.0):
    x, y = rng.uniform(-spread, spread, (2, N))
    return na.Cartesian3dVectorArray(na.ScalarArray(x, AX), na.ScalarArray(y, AX), 0.0), x, y


def rays(z=-40.0):
    pos, x, y = points()
    d = np.stack([rng.uniform(-0.05, 0.05, N), rng.uniform(-0.05, 0.05, N), np.ones(N)])
    d /= np.linalg.norm(d, axis=0)
    r = optika.rays.RayVectorArray(
        wavelength=400 * u.nm,
        position=na.Cartesian3dVectorArray(pos.x, pos.y, z),
        direction=na.Cartesian3dVectorArray(*[na.ScalarArray(c, AX) for c in d]),
        attenuation=na.ScalarArray(rng.uniform(0, 0.02, N), AX),
    )
    return r, configs.flatten_rays(r)[0]


T = tf.TransformationList([tf.Cartesian3dTranslation(x=1.0, y=-2.0), tf.Cartesian3dRotationZ(25 * u.deg)])
SAGS = [
    optika.sags.NoSag(),
    optika.sags.SphericalSag(radius=150.0),
    optika.sags.SphericalSag(radius=-150.0, transformation=T),
    optika.sags.CylindricalSag(radius=130.0),
    optika.sags.ConicSag(radius=140.0, conic=-0.6),
    optika.sags.ParabolicSag(focal_length=-70.0),
    optika.sags.ToroidalSag(radius=150.0, radius_of_rotation=170.0),
]


@pytest.mark.parametrize("sag", SAGS, ids=lambda s: type(s).__name__)
def test_sag_call_normal_intercept(sag, cuda_device):
    pos, x, y = points()
    z = sag(pos)
    assert np.allclose(z.ndarray, ora.sag_value(sag, x, y), rtol=1e-12, atol=1e-12)
    n = sag.normal(pos)
    want = ora.sag_normal(sag, x, y)
    for got, w in zip((n.x, n.y, n.z), want):
        assert np.allclose(np.broadcast_to(got.ndarray, w.shape), w, rtol=1e-12, atol=1e-14)
    r, r0 = rays()
    hit = sag.intercept(r)
    w = ora.sag_intercept(sag, r0, converge=True, extended=True)
    for got, name in ((hit.position.x, "px"), (hit.position.y, "py"), (hit.position.z, "pz")):
        assert np.allclose(got.ndarray, w[name], rtol=0, atol=1e-9 * 50)
    assert np.array_equal(hit.direction.x.ndarray, r0["dx"])  # the direction is untouched
    assert np.array_equal(hit.intensity.ndarray, r0["intensity"])
    att = sag.propagate_rays(r)
    w = ora.sag_propagate(sag, r0, converge=True, extended=True)
    assert np.allclose(att.intensity.ndarray, w["intensity"], rtol=1e-12)
    assert (att.intensity.ndarray < 1).all()


@pytest.mark.parametrize(
    "aperture",
    [
        optika.apertures.CircularAperture(12.0),
        optika.apertures.RectangularAperture(na.Cartesian2dVectorArray(15.0, 6.0), transformation=T),
        optika.apertures.OctagonalAperture(13.0, inverted=True),
        optika.apertures.EllipticalAperture(na.Cartesian2dVectorArray(15.0, 6.0)),
    ],
    ids=lambda a: type(a).__name__,
)
def test_aperture_call_and_clip(aperture, cuda_device):
    pos, x, y = points()
    mask = aperture(pos)
    assert np.array_equal(mask.ndarray, ora.aperture_mask(aperture, x, y))
    r, r0 = rays()
    clipped = aperture.clip_rays(r)
    want = ora.aperture_clip(aperture, r0)
    assert np.array_equal(clipped.unvignetted.ndarray, want["unvignetted"])
    assert np.array_equal(clipped.position.x.ndarray, r0["px"])  # only the mask changes
    assert np.array_equal(clipped.direction.z.ndarray, r0["dz"])


def unit_normals():
    nrm = np.stack([rng.uniform(-0.2, 0.2, N), rng.uniform(-0.2, 0.2, N), -np.ones(N)])
    nrm /= np.linalg.norm(nrm, axis=0)
    return na.Cartesian3dVectorArray(*[na.ScalarArray(c, AX) for c in nrm]), nrm


SPACINGS = [
    optika.rulings.ConstantRulingSpacing(constant=(1 / 1200) * u.mm, normal=na.Cartesian3dVectorArray(0.6, 0.8, 0.0)),
    optika.rulings.Polynomial1dRulingSpacing(
        coefficients={0: (1 / 1200) * u.mm, 1: 2e-8, 2: 3e-10}, normal=na.Cartesian3dVectorArray(1, 0, 0)
    ),
    optika.rulings.HolographicRulingSpacing(
        x1=na.Cartesian3dVectorArray(30.0, 0.0, -400.0), x2=na.Cartesian3dVectorArray(-50.0, 5.0, -450.0),
        wavelength=500 * u.nm,
    ),
]


@pytest.mark.parametrize("spacing", SPACINGS, ids=lambda s: type(s).__name__)
def test_ruling_spacing_and_incident_effective_with_given_normal(spacing, cuda_device):
    pos, x, y = points()
    normal, nrm = unit_normals()
    kappa = spacing(pos, normal)
    want = ora.ruling_vector(spacing, (x, y, np.zeros(N)), tuple(nrm))
    for got, w in zip((kappa.x, kappa.y, kappa.z), want):
        assert np.allclose(np.broadcast_to(got.ndarray, (N,)), w, rtol=1e-11, atol=1e-18)
    rulings = optika.rulings.Rulings(spacing=spacing, diffraction_order=-1)
    r, r0 = rays(z=0.0)
    eff = rulings.incident_effective(r, normal)
    want = ora.incident_effective(rulings, r0, tuple(nrm))
    for got, name in ((eff.direction.x, "dx"), (eff.direction.y, "dy"), (eff.direction.z, "dz")):
        assert np.allclose(got.ndarray, want[name], rtol=1e-11, atol=1e-13)
    assert not np.allclose(eff.direction.x.ndarray, r0["dx"])  # optika/rulings/_rulings_test.py:38-60


@pytest.mark.parametrize("n1,n2", [(1.0, 1.5), (1.5, 1.0), (1.2, 1.2)])
@pytest.mark.parametrize("is_mirror", [False, True])
def test_snells_law(n1, n2, is_mirror, cuda_device):
    _, r0 = rays()
    normal, nrm = unit_normals()
    direction = na.Cartesian3dVectorArray(*[na.ScalarArray(r0[k], AX) for k in ("dx", "dy", "dz")])
    got = optika.materials.snells_law(direction, n1, n2, normal=normal, is_mirror=is_mirror)
    want = ora.snells_law(r0["dx"], r0["dy"], r0["dz"], n1, n2, *nrm, is_mirror)
    for g, w in zip((got.x, got.y, got.z), want):
        assert np.allclose(g.ndarray, w, rtol=1e-12, atol=1e-14)
    assert np.allclose(got.length.ndarray, 1.0)  # optika/materials/_tests/test_snells_law.py:113
    an_in = r0["dx"] * nrm[0] + r0["dy"] * nrm[1] + r0["dz"] * nrm[2]
    an_out = got.x.ndarray * nrm[0] + got.y.ndarray * nrm[1] + got.z.ndarray * nrm[2]
    assert np.all(np.sign(an_out) == (-np.sign(an_in) if is_mirror else np.sign(an_in)))  # :114-117
    # default normal (0, 0, -1): _snells_law.py:268-269
    got0 = optika.materials.snells_law(direction, n1, n2, is_mirror=is_mirror)
    want0 = ora.snells_law(r0["dx"], r0["dy"], r0["dz"], n1, n2, 0.0, 0.0, -1.0, is_mirror)
    assert np.allclose(got0.z.ndarray, want0[2], rtol=1e-10)


def test_snells_law_with_an_index_axis(cuda_device):
    # optika/materials/_tests/test_snells_law.py:62-67: index_refraction_new on its own named axis
    _, r0 = rays()
    direction = na.Cartesian3dVectorArray(*[na.ScalarArray(r0[k], AX) for k in ("dx", "dy", "dz")])
    n2 = na.linspace(1, 2, axis="index_refraction_new", num=4)
    got = optika.materials.snells_law(direction, 1, n2)
    assert got.shape == {"index_refraction_new": 4, "ray": N}
    for i, v in enumerate(n2.ndarray):
        want = ora.snells_law(r0["dx"], r0["dy"], r0["dz"], 1.0, v, 0.0, 0.0, -1.0, False)
        assert np.allclose(got.z.numpy(("index_refraction_new", "ray"))[i], want[2], rtol=1e-12)
