"""
Pins the NumPy oracle (``oracle/raytrace.py``) to the reference's own
known-answer tests, restated here because the reference cannot be imported in
the build container.  Each test cites the reference test
it restates.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import transformations as tf
from oracle import raytrace as ora

rng = np.random.default_rng(0)


# ---------------------------------------------------------------------------
# optika/_util_test.py:20-48
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ax,ay", [(1 * u.deg, 2 * u.deg), (0.3, -0.2), (0.0, 0.0)])
def test_direction_convention(ax, ay):
    # result_expected = R_y(-ax) @ R_x(+ay) @ z-hat with right-handed matrices
    ry = ora._rotation("Y", -ax)
    rx = ora._rotation("X", +ay)
    expected = ry @ rx @ np.array([0.0, 0.0, 1.0])
    result = np.array(ora.direction(ax, ay))
    assert np.allclose(result, expected)
    # round trip, _util_test.py:36-48
    d = np.array([1.0, 2.0, 5.0])
    d /= np.linalg.norm(d)
    assert np.allclose(ora.direction(*ora.angles(*d)), d)


# ---------------------------------------------------------------------------
# optika/materials/_tests/test_snells_law.py:82-120
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n1,n2", [(1.0, 1.5), (1.5, 1.0), (1.0, 1.0), (1.2, 2.0)])
@pytest.mark.parametrize("mirror", [False, True])
def test_snells_law_vector(n1, n2, mirror):
    a = np.array([0.1, -0.2, 0.9])
    a /= np.linalg.norm(a)
    normal = np.array([0.0, 0.0, -1.0])
    b = np.array(ora.snells_law(*a, n1, n2, *normal, mirror))
    # closed formula of the reference test (test_snells_law.py:101-110)
    r = n1 / n2
    c = -a @ normal
    sign = np.sign(c) * (2 * mirror - 1)
    t = np.sqrt(np.square(1 / r) + np.square(c) - a @ a)
    expected = r * (a + (c + sign * t) * normal)
    assert np.allclose(b, expected)
    assert np.isclose(np.linalg.norm(b), 1.0)  # test_snells_law.py:113
    if mirror:  # test_snells_law.py:114-117
        assert np.sign(b @ normal) != np.sign(a @ normal)
    else:
        assert np.sign(b @ normal) == np.sign(a @ normal)


def test_snells_law_mirror_flips_normal_component():
    # test_snells_law.py:112-120: reflection == transmission with the normal component negated
    a = np.array([0.3, 0.1, 0.8])
    a /= np.linalg.norm(a)
    n = (0.0, 0.0, -1.0)
    t = np.array(ora.snells_law(*a, 1.0, 1.0, *n, False))
    r = np.array(ora.snells_law(*a, 1.0, 1.0, *n, True))
    assert np.allclose(t, a)
    assert np.allclose(r, a * np.array([1, 1, -1]))


# ---------------------------------------------------------------------------
# optika/sags/_tests/_abc_test.py:78-102 and the per-sag parametrisations
# ---------------------------------------------------------------------------
def _test_rays(n=200, spread=20.0):
    px = rng.uniform(-spread, spread, n)
    py = rng.uniform(-spread, spread, n)
    d = np.stack([rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n), np.ones(n)])
    d /= np.linalg.norm(d, axis=0)
    return ora.make_rays(n, px=px, py=py, pz=-50.0, dx=d[0], dy=d[1], dz=d[2], wavelength=5e-4)


TRANSFORMS = [
    None,
    tf.Cartesian3dTranslation(x=5 * u.mm),
    tf.TransformationList(
        [
            tf.Cartesian3dTranslation(x=5 * u.mm),
            tf.Cartesian3dRotationZ(53 * u.deg),
            tf.Cartesian3dTranslation(x=6 * u.mm),
        ]
    ),
]

SAGS = [
    lambda t: optika.sags.NoSag(transformation=t),
    lambda t: optika.sags.SphericalSag(radius=100 * u.mm, transformation=t),
    lambda t: optika.sags.SphericalSag(radius=-100 * u.mm, transformation=t),
    lambda t: optika.sags.SphericalSag(radius=1000 * u.mm, transformation=t),
    lambda t: optika.sags.CylindricalSag(radius=100 * u.mm, transformation=t),
    lambda t: optika.sags.CylindricalSag(radius=-100 * u.mm, transformation=t),
    lambda t: optika.sags.ConicSag(radius=100 * u.mm, conic=0, transformation=t),
    lambda t: optika.sags.ConicSag(radius=-100 * u.mm, conic=-1.5, transformation=t),
    lambda t: optika.sags.ConicSag(radius=100 * u.mm, conic=0.5, transformation=t),
    lambda t: optika.sags.ParabolicSag(focal_length=100 * u.mm, transformation=t),
    lambda t: optika.sags.ParabolicSag(focal_length=-100 * u.mm, transformation=t),
    lambda t: optika.sags.ToroidalSag(radius=100 * u.mm, radius_of_rotation=120 * u.mm, transformation=t),
]


@pytest.mark.parametrize("make", SAGS)
@pytest.mark.parametrize("transformation", TRANSFORMS)
def test_sag_intercept_on_surface_and_closed_form_equals_iterative(make, transformation):
    sag = make(transformation)
    rays = _test_rays()
    result = ora.sag_intercept(sag, rays)
    # _abc_test.py:98: sag(result.position) == result.position.z
    if transformation is None or type(sag).__name__ != "ToroidalSag":
        z = ora.sag_value(sag, result["px"], result["py"], result["pz"])
        if transformation is None:
            assert np.allclose(z, result["pz"])
    # _abc_test.py:100-102: closed form == AbstractSag.intercept (generic secant)
    if transformation is None:
        generic = ora.sag_intercept(sag, rays, generic=True)
        for k in ("px", "py", "pz"):
            assert np.allclose(result[k], generic[k])


@pytest.mark.parametrize("make", SAGS)
def test_sag_normal_unit_length_and_negative_z(make):
    # _abc_test.py:78-88
    sag = make(None)
    x = rng.uniform(-20, 20, 100)
    y = rng.uniform(-20, 20, 100)
    nx, ny, nz = ora.sag_normal(sag, x, y)
    assert np.all(nz < 0)
    assert np.allclose(np.sqrt(nx**2 + ny**2 + nz**2), 1)


def test_parabolic_normal_equals_conic_normal():
    # optika/sags/_tests/_parabolic_test.py:25-37
    f = 75.0
    x = rng.uniform(-30, 30, 100)
    y = rng.uniform(-30, 30, 100)
    a = ora.sag_normal(optika.sags.ParabolicSag(focal_length=f), x, y)
    b = ora.sag_normal(optika.sags.ConicSag(radius=2 * f, conic=-1), x, y)
    for p, q in zip(a, b):
        assert np.allclose(p, q)


def test_parabolic_intercept_equals_conic_intercept():
    rays = _test_rays()
    a = ora.sag_intercept(optika.sags.ParabolicSag(focal_length=-80.0), rays)
    b = ora.sag_intercept(optika.sags.ConicSag(radius=-160.0, conic=-1), rays)
    for k in ("px", "py", "pz"):
        assert np.allclose(a[k], b[k])


def test_intercept_grazing_conic():
    # optika/sags/_tests/_conic_test.py:57-87
    sag = optika.sags.ConicSag(radius=-0.7 * u.mm, conic=-1.0007)
    azimuth = np.linspace(0, 2 * np.pi, 64, endpoint=False)
    radius = 68.0
    rays = ora.make_rays(
        64, px=radius * np.cos(azimuth), py=radius * np.sin(azimuth), pz=-2000.0, dz=1.0
    )
    p = ora.sag_intercept(sag, rays)
    assert np.allclose(ora.sag_value(sag, p["px"], p["py"]), p["pz"])
    c = 1 / -0.7
    r2 = p["px"] ** 2 + p["py"] ** 2
    assert np.all((p["pz"] * (c * r2 - p["pz"])) >= 0)


# ---------------------------------------------------------------------------
# optika/materials/_tests/test_materials.py:126-150
# ---------------------------------------------------------------------------
@pytest.mark.parametrize(
    "glass,n_d", [(optika.materials.Glass.n_bk7(), 1.5168), (optika.materials.Glass.f2(), 1.6200)]
)
def test_glass_index(glass, n_d):
    rays = ora.make_rays(1, wavelength=587.56 * u.nm)
    assert abs(ora.index_refraction(glass, rays)[0] - n_d) < 1e-3


# ---------------------------------------------------------------------------
# optika/rulings/_spacing.py:141-219 (the holographic refocusing example):
# rays from x1 diffract towards x2 at the recording wavelength
# ---------------------------------------------------------------------------
def test_holographic_rulings_refocus():
    x1 = na.Cartesian3dVectorArray(10.0, 0.0, -100.0)
    x2 = na.Cartesian3dVectorArray(-20.0, 0.0, -150.0)
    w = 500 * u.nm
    spacing = optika.rulings.HolographicRulingSpacing(
        x1=x1, x2=x2, wavelength=w, is_diverging_1=True, is_diverging_2=False
    )
    rulings = optika.rulings.Rulings(spacing=spacing, diffraction_order=1)
    surface = optika.surfaces.Surface(rulings=rulings, material=optika.materials.Mirror())
    # rays diverging from x1 towards points on the flat grating
    n = 50
    tx = rng.uniform(-5, 5, n)
    ty = rng.uniform(-5, 5, n)
    d = np.stack([tx - 10.0, ty - 0.0, 0.0 - (-100.0) + 0 * tx])
    d /= np.linalg.norm(d, axis=0)
    rays = ora.make_rays(n, px=10.0, py=0.0, pz=-100.0, dx=d[0], dy=d[1], dz=d[2], wavelength=w)
    best = None
    for order in (1, -1):
        rulings.diffraction_order = order
        out = ora.surface_propagate(surface, rays)
        # distance of closest approach of each output ray to x2
        p = np.stack([out["px"], out["py"], out["pz"]])
        v = np.stack([out["dx"], out["dy"], out["dz"]])
        to = np.array([-20.0, 0.0, -150.0])[:, None] - p
        miss = np.linalg.norm(to - (np.sum(to * v, axis=0) / np.sum(v * v, axis=0)) * v, axis=0)
        best = miss.max() if best is None else min(best, miss.max())
    assert best < 1e-9


# ---------------------------------------------------------------------------
# apertures: optika/apertures/_apertures_test.py:62-96, 344-373
# ---------------------------------------------------------------------------
def test_rectangular_aperture_decentred():
    ap = optika.apertures.RectangularAperture(
        half_width=na.Cartesian2dVectorArray(10.0, 5.0),
        transformation=tf.Cartesian3dTranslation(x=20 * u.mm),
    )
    assert ora.aperture_mask(ap, 20.0, 0.0)
    assert ora.aperture_mask(ap, 29.0, 4.0)
    assert not ora.aperture_mask(ap, 0.0, 0.0)
    assert not ora.aperture_mask(ap, 31.0, 0.0)


@pytest.mark.parametrize(
    "aperture",
    [
        optika.apertures.CircularAperture(10.0),
        optika.apertures.RectangularAperture(10.0),
        optika.apertures.OctagonalAperture(10.0),
        optika.apertures.RegularPolygonalAperture(10.0, 6),
        optika.apertures.EllipticalAperture(na.Cartesian2dVectorArray(10.0, 5.0)),
        optika.apertures.IsoscelesTrapezoidalAperture(x_left=2.0, x_right=10.0, angle=45 * u.deg),
    ],
)
def test_aperture_semantics(aperture):
    import dataclasses

    x = rng.uniform(-15, 15, 500)
    y = rng.uniform(-15, 15, 500)
    mask = ora.aperture_mask(aperture, x, y)
    assert mask.any() and not mask.all()
    # inactive => all True (_apertures_test.py:62-72)
    assert ora.aperture_mask(dataclasses.replace(aperture, active=False), x, y).all()
    # inverted => complement
    assert np.array_equal(ora.aperture_mask(dataclasses.replace(aperture, inverted=True), x, y), ~mask)
    # clip only changes the mask (_apertures_test.py:82-96)
    rays = ora.make_rays(500, px=x, py=y)
    clipped = ora.aperture_clip(aperture, rays)
    for k in ora.FIELDS:
        assert np.array_equal(clipped[k], rays[k])
    assert np.array_equal(clipped["unvignetted"], mask)


def test_polygon_matches_rectangle_and_circle_limit():
    # a rectangle tested as a generic polygon agrees with the rectangular rule away from edges
    rect = optika.apertures.RectangularAperture(na.Cartesian2dVectorArray(7.0, 3.0))
    poly = optika.apertures.PolygonalAperture(vertices=rect.vertices)
    x = rng.uniform(-10, 10, 2000)
    y = rng.uniform(-10, 10, 2000)
    far = ora.aperture_margin(rect, x, y) > 1e-9
    assert np.array_equal(ora.aperture_mask(rect, x, y)[far], ora.aperture_mask(poly, x, y)[far])


# ---------------------------------------------------------------------------
# end to end: the Newtonian example focuses (geometry pins the transformation
# conventions: optika/systems/_sequential.py:1882-1912)
# ---------------------------------------------------------------------------
def test_newtonian_focus_and_conventions():
    import configs

    system = configs.newtonian(num_field=1, num_pupil=8)
    _, rays = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    out = ora.propagate_rays(system.surfaces_all, r0)
    local = ora._rays_transform(system.sensor.transformation, out, inverse=True)
    m = out["unvignetted"]
    assert m.sum() > 10
    # on-axis field of a paraboloid: a perfect focus at the sensor centre
    assert np.max(np.abs(local["px"][m])) < 1e-9
    assert np.max(np.abs(local["py"][m])) < 1e-9
    assert np.max(np.abs(local["pz"])) < 1e-9
    # rays arrive travelling along +z of the sensor frame
    assert np.all(local["dz"][m] > 0.9)


def test_numba_snell_twin_is_bit_identical_to_the_numpy_expression():
    """
    The CPU arm runs Snell's law as the reference does, through a numba ``guvectorize`` kernel
    (``optika/materials/_snells_law.py:294-366``); the oracle's NumPy expression and its numba twin
    (parallel and serial targets) must agree to the last bit, NaN patterns included.
    """
    pytest.importorskip("numba")
    from oracle import snell_numba

    rng = np.random.default_rng(5)
    n = 20000
    a = rng.normal(size=(3, n))
    a /= np.linalg.norm(a, axis=0) * rng.uniform(0.8, 1.2, n)  # not unit: |a|^2 enters the formula
    normal = rng.normal(size=(3, n))
    normal /= np.linalg.norm(normal, axis=0)
    n1, n2 = rng.uniform(1, 1.7, n), rng.uniform(1, 1.7, n)
    for mirror in (False, True):
        want = ora.snells_law(*a, n1, n2, *normal, mirror)
        assert np.isnan(want[0]).any() or mirror  # total internal reflection occurs in the sample
        for kernel in (snell_numba.snells_law_parallel, snell_numba.snells_law_serial):
            with np.errstate(invalid="ignore"):
                got = kernel(*a, n1, n2, *normal, mirror)
            for g, w in zip(got, want):
                assert np.array_equal(g, w, equal_nan=True)
    try:
        assert ora.use_numba_snell("parallel") == "parallel"
        through_switch = ora.snells_law(*a, n1, n2, *normal, True)
    finally:
        ora.use_numba_snell(None)
    assert all(np.array_equal(g, w, equal_nan=True) for g, w in zip(through_switch, ora.snells_law(*a, n1, n2, *normal, True)))
