"""
Parity of the fused CUDA trace (through the C ABI) against the NumPy oracle on
identical inputs.  North-star tolerances: positions / directions within 1e-9
relative (fp64); ``unvignetted`` bit-exact except enumerated aperture-edge rays.
"""

import ctypes as C
import dataclasses

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import transformations as tf
from optika_b200 import _engine, _lib
from oracle import raytrace as ora

import configs
import parity

pytestmark = pytest.mark.gpu

rng = np.random.default_rng(1)


def host_states(rays_host, axis=None) -> dict:
    """RayVectorArray -> dict of arrays [n_surface, n_rays] (or [n_rays])."""
    d, shape_ = configs.flatten_rays(rays_host)
    if axis is not None:
        k = list(shape_).index(axis)
        d = {name: np.moveaxis(v, k, 0).reshape(shape_[axis], -1) for name, v in d.items()}
    else:
        d = {name: v.reshape(-1) for name, v in d.items()}
    return d


def check_system(system, cuda_device, accumulate=True):
    _, rays = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    surfaces = system.surfaces_all
    if accumulate:
        dev = optika.propagators.accumulate_rays(surfaces, rays, axis="surface")
        got = host_states(dev, axis="surface")
        want = ora.accumulate_rays(surfaces, r0, converge=True, extended=True)
    else:
        dev = optika.propagators.propagate_rays(surfaces, rays)
        got = host_states(dev)
        want = ora.propagate_rays(surfaces, r0, converge=True, extended=True)
    return parity.compare_states(got, want, surfaces if accumulate else None)


def test_cfg1_newtonian_every_ray_every_surface(cuda_device):
    # BASELINE config 1: 10 x 10 field x 32 x 32 pupil = 102 400 rays, 6 surfaces
    system = configs.newtonian(num_field=10, num_pupil=32)
    report = check_system(system, cuda_device, accumulate=True)
    assert report["position"] <= parity.RTOL
    assert report["mask_mismatches"] == 0


def test_cfg1_float64_noise_of_the_reference_formula_is_enumerated(cuda_device):
    """
    The reference's parabolic closed form (``optika/sags/_parabolic.py:142-151``)
    cancels for near-axial rays: evaluated in float64 it is ~1e-6 mm away from
    its own exact value for the smallest field angles of cfg 1.  Rays for which
    the float64 and the extended evaluation of the SAME formula differ by more
    than 1e-10 relative are enumerated; every other ray agrees with the float64
    oracle within 1e-9, and the enumerated ones within the reference's own noise.
    """
    system = configs.newtonian(num_field=10, num_pupil=32)
    _, rays = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    surfaces = system.surfaces_all
    got = host_states(optika.propagators.accumulate_rays(surfaces, rays, axis="surface"), axis="surface")
    w64 = ora.accumulate_rays(surfaces, r0)
    w80 = ora.accumulate_rays(surfaces, r0, extended=True)
    n_noisy = 0
    for s in range(len(surfaces)):
        scale = max(np.abs(w64[k][s]).max() for k in ("px", "py", "pz"))
        noise = np.sqrt(sum((w64[k][s] - w80[k][s]) ** 2 for k in ("px", "py", "pz")))
        err = np.sqrt(sum((got[k][s] - w64[k][s]) ** 2 for k in ("px", "py", "pz")))
        noisy = noise > 1e-10 * scale
        n_noisy += int(noisy.sum())
        assert np.all(err[~noisy] <= 1e-9 * scale)
        assert np.all(err[noisy] <= 2 * noise[noisy] + 1e-9 * scale)
    # the enumerated set is the small-field-angle rays after the primary mirror (a few percent)
    assert 0 < n_noisy < 0.3 * got["px"].size


def test_cfg1_raytrace_api_matches_propagators(cuda_device):
    system = configs.newtonian(num_field=3, num_pupil=8)
    a = system.raytrace(accumulate=True, **configs.PHYSICAL).outputs
    b = optika.propagators.accumulate_rays(system.surfaces_all, system._input(None, None, None, None, False, False)[1], axis="surface")
    for x, y in ((a.position.x, b.position.x), (a.direction.z, b.direction.z)):
        # named axes: raytrace() orders the device grid (field outer, pupil inner), values are identical
        assert np.array_equal(x.numpy(tuple(y.shape)), y.ndarray)
    assert "surface" in a.shape and a.shape["surface"] == 6
    # rayfunction: last surface, sensor-local coordinates (optika/systems/_sequential.py:970-986)
    local = system.rayfunction(**configs.PHYSICAL).outputs
    last = {k: v[-1] for k, v in host_states(a, axis="surface").items()}
    want = ora._rays_transform(system.sensor.transformation, last, inverse=True)
    got = host_states(local)
    assert np.allclose(got["px"], want["px"], rtol=0, atol=1e-9 * 50)
    assert np.allclose(got["dz"], want["dz"], rtol=0, atol=1e-12)
    assert np.max(np.abs(got["pz"])) < 1e-9 * 50


def test_cfg2_spherical_grating(cuda_device):
    system = configs.spherical_grating(num_field=6, num_pupil=16, num_wavelength=5)
    report = check_system(system, cuda_device, accumulate=True)
    assert report["direction"] <= parity.RTOL


def test_cfg3_toroidal_vls_octagon(cuda_device):
    system = configs.toroidal_vls(num_field=5, num_pupil=14, num_wavelength=3)
    report = check_system(system, cuda_device, accumulate=True)
    assert report["position"] <= parity.RTOL


def test_cfg3_toroid_matches_reference_secant_within_its_truncation(cuda_device):
    """
    The reference's generic intercept stops its secant at |step| < 1e-6 mm
    (``optika/sags/_abc.py:99-103``); the device iterates Newton to convergence.
    The two agree to far better than the north-star tolerance.
    """
    system = configs.toroidal_vls(num_field=3, num_pupil=10, num_wavelength=1)
    _, rays = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    got = host_states(optika.propagators.propagate_rays(system.surfaces_all, rays))
    want = ora.propagate_rays(system.surfaces_all, r0, converge=False)
    parity.compare_states(got, want)


def test_cfg5_configuration_axis(cuda_device):
    system = configs.misaligned_telescope(num_field=3, num_pupil=10, num_tilt=4)
    assert system.shape == {"misalign": 4}
    _, rays = system._input(None, None, None, None, False, False)
    dev = optika.propagators.accumulate_rays(system.surfaces_all, rays, axis="surface")
    assert list(dev.shape)[:2] == ["misalign", "surface"]
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    for i in range(4):
        surfaces = [ora.select_config(s, {"misalign": i}) for s in system.surfaces_all]
        want = ora.accumulate_rays(surfaces, r0, converge=True, extended=True)
        got = host_states(dev[{"misalign": i}], axis="surface")
        parity.compare_states(got, want, surfaces)
    # the tilt really changes the answer
    a = dev[{"misalign": 0}].position.x.ndarray
    b = dev[{"misalign": 3}].position.x.ndarray
    assert not np.allclose(a, b)


# ---------------------------------------------------------------------------
# every sag / aperture / ruling / material kind, with and without transformations
# ---------------------------------------------------------------------------
def random_rays(n=4096, spread=20.0, z=-50.0, tilt=0.05, wavelength=500 * u.nm):
    px = rng.uniform(-spread, spread, n)
    py = rng.uniform(-spread, spread, n)
    d = np.stack([rng.uniform(-tilt, tilt, n), rng.uniform(-tilt, tilt, n), np.ones(n)])
    d /= np.linalg.norm(d, axis=0)
    ax = "ray"
    rays = optika.rays.RayVectorArray(
        wavelength=wavelength,
        position=na.Cartesian3dVectorArray(na.ScalarArray(px, ax), na.ScalarArray(py, ax), z),
        direction=na.Cartesian3dVectorArray(
            na.ScalarArray(d[0], ax), na.ScalarArray(d[1], ax), na.ScalarArray(d[2], ax)
        ),
        intensity=na.ScalarArray(rng.uniform(0.5, 1.5, n), ax),
    )
    return rays


def check_surface(surface, rays, cuda_device):
    r0, _ = configs.flatten_rays(rays)
    got = host_states(surface.propagate_rays(rays))
    want = ora.surface_propagate(surface, r0, converge=True, extended=True)
    return parity.compare_states(
        {k: v[None] for k, v in got.items()}, {k: v[None] for k, v in want.items()}, [surface]
    )


T_LIST = tf.TransformationList(
    [
        tf.Cartesian3dRotationX(3 * u.deg),
        tf.Cartesian3dRotationY(-2 * u.deg),
        tf.Cartesian3dTranslation(x=1.5, y=-0.5, z=30.0),
        tf.Cartesian3dRotationZ(20 * u.deg),
    ]
)
T_SMALL = tf.TransformationList([tf.Cartesian3dTranslation(x=2.0, y=1.0), tf.Cartesian3dRotationZ(30 * u.deg)])

SAGS = [
    optika.sags.NoSag(),
    optika.sags.NoSag(transformation=T_SMALL),
    optika.sags.SphericalSag(radius=200.0),
    optika.sags.SphericalSag(radius=-150.0, transformation=T_SMALL),
    optika.sags.CylindricalSag(radius=120.0),
    optika.sags.CylindricalSag(radius=-120.0, transformation=T_SMALL),
    optika.sags.ConicSag(radius=150.0, conic=-0.5),
    optika.sags.ConicSag(radius=-150.0, conic=-2.0, transformation=T_SMALL),
    optika.sags.ConicSag(radius=100.0, conic=0.7),
    optika.sags.ParabolicSag(focal_length=90.0),
    optika.sags.ParabolicSag(focal_length=-90.0, transformation=T_SMALL),
    optika.sags.ToroidalSag(radius=150.0, radius_of_rotation=180.0),
    optika.sags.ToroidalSag(radius=-150.0, radius_of_rotation=200.0),
]


@pytest.mark.parametrize("sag", SAGS, ids=lambda s: type(s).__name__)
@pytest.mark.parametrize("material", [optika.materials.Mirror(), optika.materials.Glass.n_bk7()], ids=lambda m: type(m).__name__)
@pytest.mark.parametrize("transformation", [None, T_LIST], ids=["identity", "transformed"])
def test_all_sags(sag, material, transformation, cuda_device):
    surface = optika.surfaces.Surface(sag=sag, material=material, transformation=transformation)
    check_surface(surface, random_rays(), cuda_device)


APERTURES = [
    optika.apertures.CircularAperture(12.0),
    optika.apertures.CircularAperture(12.0, inverted=True, transformation=T_SMALL),
    optika.apertures.CircularAperture(12.0, active=False),
    optika.apertures.RectangularAperture(na.Cartesian2dVectorArray(15.0, 8.0)),
    optika.apertures.RectangularAperture(10.0, inverted=True),
    optika.apertures.EllipticalAperture(na.Cartesian2dVectorArray(15.0, 8.0), transformation=T_SMALL),
    optika.apertures.CircularSectorAperture(15.0, angle_start=10 * u.deg, angle_stop=200 * u.deg),
    optika.apertures.CircularSectorAperture(15.0, angle_start=-170 * u.deg, angle_stop=-20 * u.deg),
    optika.apertures.OctagonalAperture(14.0),
    optika.apertures.RegularPolygonalAperture(14.0, 5, transformation=T_SMALL),
    optika.apertures.RegularPolygonalAperture(14.0, 6, inverted=True),
    optika.apertures.IsoscelesTrapezoidalAperture(x_left=3.0, x_right=16.0, angle=50 * u.deg),
    optika.apertures.PolygonalAperture(
        vertices=na.Cartesian3dVectorArray(
            na.ScalarArray(np.array([-10.0, 12.0, 4.0, 9.0, -6.0]), "vertex"),
            na.ScalarArray(np.array([-8.0, -9.0, 0.0, 11.0, 7.0]), "vertex"),
            na.ScalarArray(np.zeros(5), "vertex"),
        )
    ),
    optika.apertures.PolygonalAperture(
        vertices=na.Cartesian3dVectorArray(
            na.ScalarArray(np.array([-10.0, 12.0, 4.0]), "vertex"),
            na.ScalarArray(np.array([-8.0, -9.0, 10.0]), "vertex"),
            na.ScalarArray(np.zeros(3), "vertex"),
        ),
        active=False,
    ),
]


@pytest.mark.parametrize("aperture", APERTURES, ids=lambda a: type(a).__name__)
def test_all_apertures(aperture, cuda_device):
    surface = optika.surfaces.Surface(aperture=aperture, transformation=T_LIST)
    report = check_surface(surface, random_rays(n=20000), cuda_device)
    assert report["mask_mismatches"] <= 2


def test_aperture_edges_exactly_on_grid_points(cuda_device):
    """Rays landing EXACTLY on aperture edges: inclusive comparisons must agree bit for bit."""
    x = np.linspace(-12, 12, 49)  # includes +-10 and +-5 exactly
    xx, yy = np.meshgrid(x, x, indexing="ij")
    ax = "ray"
    rays = optika.rays.RayVectorArray(
        wavelength=5e-4,
        position=na.Cartesian3dVectorArray(na.ScalarArray(xx.ravel(), ax), na.ScalarArray(yy.ravel(), ax), -1.0),
        direction=na.Cartesian3dVectorArray(0.0, 0.0, 1.0),
    )
    r0, _ = configs.flatten_rays(rays)
    for aperture in (
        optika.apertures.RectangularAperture(na.Cartesian2dVectorArray(10.0, 5.0)),
        optika.apertures.CircularAperture(10.0),
        optika.apertures.EllipticalAperture(na.Cartesian2dVectorArray(10.0, 5.0)),
        optika.apertures.RegularPolygonalAperture(10.0, 4),
    ):
        surface = optika.surfaces.Surface(aperture=aperture)
        got = host_states(surface.propagate_rays(rays))
        want = ora.surface_propagate(surface, r0)
        assert np.array_equal(got["unvignetted"], want["unvignetted"]), type(aperture).__name__


def test_angular_aperture_clips_on_direction(cuda_device):
    aperture = optika.apertures.CircularAperture(0.03, angular=True)
    surface = optika.surfaces.Surface(aperture=aperture)
    report = check_surface(surface, random_rays(n=5000), cuda_device)
    assert report["mask_mismatches"] == 0


RULINGS = [
    optika.rulings.Rulings(spacing=(1 / 600) * u.mm, diffraction_order=1),
    optika.rulings.Rulings(spacing=(1 / 600) * u.mm, diffraction_order=-2),
    optika.rulings.Rulings(
        spacing=optika.rulings.ConstantRulingSpacing(
            constant=(1 / 900) * u.mm, normal=na.Cartesian3dVectorArray(0.6, 0.8, 0.0)
        ),
        diffraction_order=1,
    ),
    optika.rulings.Rulings(
        spacing=optika.rulings.Polynomial1dRulingSpacing(
            coefficients={0: (1 / 1200) * u.mm, 1: 3e-8, 2: 2e-10, 3: -1e-12},
            normal=na.Cartesian3dVectorArray(1, 0, 0),
        ),
        diffraction_order=1,
    ),
    optika.rulings.Rulings(
        spacing=optika.rulings.Polynomial1dRulingSpacing(
            coefficients={0: (1 / 1200) * u.mm, 2: 2e-10},
            normal=na.Cartesian3dVectorArray(0, 1, 0),
            transformation=tf.Cartesian3dTranslation(y=3.0),
        ),
        diffraction_order=2,
    ),
    optika.rulings.Rulings(
        spacing=optika.rulings.HolographicRulingSpacing(
            x1=na.Cartesian3dVectorArray(30.0, 0.0, -400.0),
            x2=na.Cartesian3dVectorArray(-50.0, 5.0, -450.0),
            wavelength=500 * u.nm,
        ),
        diffraction_order=1,
    ),
    optika.rulings.Rulings(
        spacing=optika.rulings.HolographicRulingSpacing(
            x1=na.Cartesian3dVectorArray(30.0, 0.0, -400.0),
            x2=na.Cartesian3dVectorArray(-50.0, 5.0, -450.0),
            wavelength=500 * u.nm,
            is_diverging_1=False,
            is_diverging_2=True,
        ),
        diffraction_order=-1,
    ),
]


@pytest.mark.parametrize("rulings", RULINGS, ids=lambda r: type(r.spacing_).__name__)
@pytest.mark.parametrize(
    "sag", [optika.sags.NoSag(), optika.sags.SphericalSag(radius=-500.0), optika.sags.ToroidalSag(400.0, 450.0)],
    ids=lambda s: type(s).__name__,
)
def test_all_rulings(rulings, sag, cuda_device):
    surface = optika.surfaces.Surface(
        sag=sag, rulings=rulings, material=optika.materials.Mirror(), transformation=T_LIST
    )
    check_surface(surface, random_rays(wavelength=300 * u.nm), cuda_device)


def test_glass_lens_wavelength_rescale_and_attenuation(cuda_device):
    """Two glass interfaces: wavelength is rescaled by n2/n1, attenuation handled (surfaces.py:163-190)."""
    front = optika.surfaces.Surface(
        sag=optika.sags.SphericalSag(radius=80.0), material=optika.materials.Glass.n_bk7(),
        aperture=optika.apertures.CircularAperture(18.0),
    )
    back = optika.surfaces.Surface(
        sag=optika.sags.SphericalSag(radius=-80.0), material=optika.materials.Vacuum(),
        transformation=tf.Cartesian3dTranslation(z=6.0),
    )
    image = optika.surfaces.Surface(transformation=tf.Cartesian3dTranslation(z=80.0))
    rays = random_rays()
    rays.attenuation = na.ScalarArray(rng.uniform(0, 0.01, 4096), "ray")
    r0, _ = configs.flatten_rays(rays)
    surfaces = [front, back, image]
    got = host_states(optika.propagators.accumulate_rays(surfaces, rays, axis="surface"), axis="surface")
    want = ora.accumulate_rays(surfaces, r0, converge=True, extended=True)
    parity.compare_states(got, want, surfaces)
    assert not np.allclose(want["wavelength"][0], want["wavelength"][1])


def test_systems_longer_than_one_launch_are_chained(cuda_device):
    """
    A periscope of flat mirrors with more surfaces than one launch carries (OPTK_MAX_SURFACES): propagate_rays
    and accumulate_rays chain launches, every state of every surface equals the oracle's.
    """
    n = _lib.MAX_SURFACES + 7
    surfaces = []
    for k in range(n):
        surfaces.append(optika.surfaces.Surface(
            name=f"fold_{k}",
            material=optika.materials.Mirror() if k % 3 else optika.materials.Vacuum(),
            aperture=optika.apertures.CircularAperture(24.0 + k),
            transformation=tf.TransformationList([
                tf.Cartesian3dRotationY((3.0 if k % 2 else -3.0) * u.deg),
                tf.Cartesian3dTranslation(x=0.3 * k, z=(40.0 if (k // 1) % 2 == 0 else -40.0) + 0.5 * k),
            ]),
        ))
    rays = random_rays(n=3000)
    r0, _ = configs.flatten_rays(rays)
    got = host_states(optika.propagators.accumulate_rays(surfaces, rays, axis="surface"), axis="surface")
    want = ora.accumulate_rays(surfaces, r0, converge=True, extended=True)
    assert got["px"].shape[0] == n
    parity.compare_states(got, want, surfaces)
    last = host_states(optika.propagators.propagate_rays(surfaces, rays))
    for name in ("px", "py", "pz", "dx", "dy", "dz"):
        assert np.allclose(last[name], want[name][-1], rtol=0, atol=1e-9 * 100, equal_nan=True), name
    assert np.array_equal(last["unvignetted"], want["unvignetted"][-1])


def test_missed_surfaces_propagate_nan_and_inf(cuda_device):
    """Rays that miss a sphere give NaN, a conic gives inf (optika/sags/_conic.py:154), as in the reference."""
    rays = random_rays(n=2000, spread=400.0)
    for sag in (optika.sags.SphericalSag(radius=100.0), optika.sags.ConicSag(radius=100.0, conic=0.5)):
        surface = optika.surfaces.Surface(sag=sag, material=optika.materials.Mirror())
        r0, _ = configs.flatten_rays(rays)
        got = host_states(surface.propagate_rays(rays))
        want = ora.surface_propagate(surface, r0)
        assert (~np.isfinite(want["px"])).any()
        parity.compare_states({k: v[None] for k, v in got.items()}, {k: v[None] for k, v in want.items()}, [surface])


def test_empty_and_scalar_ray_sets(cuda_device):
    surface = optika.surfaces.Surface(sag=optika.sags.SphericalSag(radius=100.0))
    scalar = optika.rays.RayVectorArray(
        wavelength=5e-4,
        position=na.Cartesian3dVectorArray(1.0, 2.0, -5.0),
        direction=na.Cartesian3dVectorArray(0.0, 0.0, 1.0),
    )
    out = surface.propagate_rays(scalar)
    assert out.shape == {}
    want = ora.surface_propagate(surface, configs.flatten_rays(scalar)[0])
    assert np.isclose(float(out.position.z), float(want["pz"]), rtol=1e-12)
    empty = optika.rays.RayVectorArray(
        wavelength=5e-4,
        position=na.Cartesian3dVectorArray(na.ScalarArray(np.zeros(0), "ray"), 0.0, -5.0),
        direction=na.Cartesian3dVectorArray(0.0, 0.0, 1.0),
    )
    out = surface.propagate_rays(empty)
    assert out.shape == {"ray": 0}


def test_backwards_trace(cuda_device):
    """surf_step = -1: the stop solver traces a subsystem backwards (optika/systems/_sequential.py:640-654)."""
    system = configs.newtonian(num_field=2, num_pupil=6)
    surfaces = system.surfaces_all[:4]
    _, rays = system._input(None, None, None, None, False, False)
    compiled = _engine.CompiledSystem(surfaces)
    got = host_states(_engine.trace(compiled, rays, surf_begin=3, surf_count=4, surf_step=-1).to_host())
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: v.reshape(-1) for k, v in r0.items()}
    want = ora.propagate_rays(surfaces[::-1], r0, extended=True)
    parity.compare_states(got, want)


def test_dense_and_broadcast_inputs_agree(cuda_device):
    """The separable (stride-0) input path and the dense path are the same rays."""
    system = configs.spherical_grating(num_field=4, num_pupil=8, num_wavelength=3)
    _, rays = system._input(None, None, None, None, False, False)
    a = _engine.trace(system._compiled, rays)
    dense = optika.sensors._host_to_device(rays, cuda_device)
    b = _engine.trace(system._compiled, dense)
    for k in a.fields:
        assert np.array_equal(a.fields[k].cpu().numpy(), b.fields[k].cpu().numpy(), equal_nan=True)
    assert np.array_equal(a.unvignetted.cpu().numpy(), b.unvignetted.cpu().numpy())


def test_host_pointer_entry_point(cuda_device):
    """optk_trace_host: host buffers in, host buffers out, streamed in slabs (pinned and pageable)."""
    system = configs.newtonian(num_field=4, num_pupil=16)
    _, rays = system._input(None, None, None, None, False, False)
    r0, shape_ = configs.flatten_rays(rays)
    n = r0["px"].size
    compiled = system._compiled
    lib = _lib.lib()
    n_surf = compiled.n_surface
    for accumulate in (0, 1):
        for slab in (0, 1000):
            rin = _lib.RaysIn()
            rin.n_axes = 1
            rin.dims[0] = n
            arrays = [np.ascontiguousarray(r0[name].reshape(-1)) for name in _lib.FIELDS]
            for f, a in enumerate(arrays):
                rin.field[f] = a.ctypes.data
                rin.stride[f][0] = 1
            rin.unvignetted = None
            states = n_surf if accumulate else 1
            outs = [np.full(states * n, np.nan) for _ in _lib.FIELDS]
            mask = np.zeros(states * n, dtype=np.uint8)
            rout = _lib.RaysOut()
            for f, a in enumerate(outs):
                rout.field[f] = a.ctypes.data
            rout.unvignetted = mask.ctypes.data
            stats = _lib.TraceStats()
            _lib.check(
                lib.optk_trace_host(
                    compiled.handle, 0, C.byref(rin), C.byref(rout), 0, n_surf, 1, accumulate, n,
                    None, None, C.byref(stats), slab, 0,
                )
            )
            assert stats.n_rays == n
            want = ora.accumulate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
            got = {name: a.reshape(states, n) for name, a in zip(_lib.FIELDS, outs)}
            got["unvignetted"] = mask.reshape(states, n).astype(bool)
            if not accumulate:
                want = {k: v[-1:] for k, v in want.items()}
            parity.compare_states(got, want, system.surfaces_all if accumulate else None)
            assert stats.n_unvignetted == int(want["unvignetted"][-1].sum())


def test_unsupported_and_invalid_arguments_raise(cuda_device):
    class WeirdSag(optika.sags.AbstractSag):
        transformation = None

    with pytest.raises(NotImplementedError):
        _engine.CompiledSystem([optika.surfaces.Surface(sag=WeirdSag())])
    system = configs.newtonian(num_field=1, num_pupil=2)
    with pytest.raises(ValueError):
        _engine.trace(system._compiled, system._input(None, None, None, None, False, False)[1], surf_begin=4, surf_count=6)


def test_host_pointer_entry_point_broadcast_inputs_and_image(cuda_device):
    """
    optk_trace_host with a separable (stride-0) host grid, several slabs, ray outputs AND a
    fused detector image, all in host memory.
    """
    from oracle import binning as orb

    system = configs.newtonian(num_field=3, num_pupil=24, num_pixel=64)
    _, rays = system._input(None, None, None, None, False, False)
    r0, shape_ = configs.flatten_rays(rays)  # axes: pupil_x, pupil_y, field_y, field_x
    assert list(shape_) == ["pupil_x", "pupil_y", "field_y", "field_x"]
    n = int(np.prod(list(shape_.values())))
    npx, npy, nfy, nfx = shape_.values()
    compiled = system._compiled_local
    lib = _lib.lib()
    small = {
        "wavelength": (np.array([float(rays.wavelength)]), (0, 0, 0, 0)),
        "px": (np.ascontiguousarray(rays.position.x.ndarray), (1, 0, 0, 0)),
        "py": (np.ascontiguousarray(rays.position.y.ndarray), (0, 1, 0, 0)),
        "pz": (np.array([0.0]), (0, 0, 0, 0)),
        "intensity": (np.array([1.0]), (0, 0, 0, 0)),
        "attenuation": (np.array([0.0]), (0, 0, 0, 0)),
        "index_refraction": (np.array([1.0]), (0, 0, 0, 0)),
    }
    for name, comp in (("dx", rays.direction.x), ("dy", rays.direction.y), ("dz", rays.direction.z)):
        nd = np.ascontiguousarray(comp.numpy(("field_y", "field_x")).astype(float) + np.zeros((nfy, nfx)))
        small[name] = (nd, (0, 0, nfx, 1))
    rin = _lib.RaysIn()
    rin.n_axes = 4
    for a, d in enumerate((npx, npy, nfy, nfx)):
        rin.dims[a] = d
    for f, name in enumerate(_lib.FIELDS):
        nd, strides = small[name]
        rin.field[f] = nd.ctypes.data
        for a, st in enumerate(strides):
            rin.stride[f][a] = st
    rin.unvignetted = None
    outs = [np.full(n, np.nan) for _ in _lib.FIELDS]
    mask = np.zeros(n, dtype=np.uint8)
    rout = _lib.RaysOut()
    for f, a in enumerate(outs):
        rout.field[f] = a.ctypes.data
    rout.unvignetted = mask.ctypes.data
    ex, ey = system.sensor.pixel_edges()
    ew = np.array([4.99e-4, 5.01e-4])
    flux = np.zeros((1, 64, 64))
    counts = np.zeros((1, 64, 64), dtype=np.uint64)
    image = _lib.Image()
    image.n_wavelength, image.n_x, image.n_y = 1, 64, 64
    image.edges_wavelength, image.edges_x, image.edges_y = ew.ctypes.data, ex.ctypes.data, ey.ctypes.data
    image.flux, image.counts = flux.ctypes.data, counts.ctypes.data
    stats = _lib.TraceStats()
    _lib.check(
        lib.optk_trace_host(
            compiled.handle, 0, C.byref(rin), C.byref(rout), 0, compiled.n_surface, 1, 0, 0,
            C.byref(image), None, C.byref(stats), 700, 0,
        )
    )
    want = ora.propagate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
    local = ora._rays_transform(system.sensor.transformation, want, inverse=True)
    got = {name: a for name, a in zip(_lib.FIELDS, outs)}
    got["unvignetted"] = mask.astype(bool)
    parity.compare_states(got, local)
    want_counts = orb.counts(local, ew, ex, ey)
    assert counts.sum() == want_counts.sum() == stats.n_unvignetted
    assert (counts.astype(np.int64) != want_counts).sum() <= 4
    assert np.isclose(flux.sum(), want_counts.sum())


def test_streaming_pipeline_kernel_is_bit_identical(cuda_device):
    """
    Dense device rays (>= 64 tiles of 512) take the persistent bulk-copy pipeline
    (trace_tma.cu) plus the ordinary kernel for the remainder; broadcast host grids take the
    ordinary kernels.  Same arithmetic, so the results must agree bit for bit.
    """
    from optika_b200 import sensors

    system = configs.spherical_grating(num_field=40, num_pupil=9, num_wavelength=1, num_pixel=256)
    _, rays = system._input(None, None, None, None, False, False)
    order = system._ray_axes_order
    compiled = system._compiled
    want, want_stats = _engine.trace(compiled, rays, ray_axes_order=order, stats=True)
    n = want.size
    assert n == 129600 and n % 512 == 64  # 253 full tiles and a tail
    # dense input in the same ray order as `want`
    first = _engine.trace(compiled, rays, ray_axes_order=order, surf_count=0)
    got, got_stats = _engine.trace(compiled, first, stats=True)
    assert got_stats == want_stats
    for name in want.fields:
        assert torch_equal(got.fields[name], want.fields[name]), name
    assert torch_equal(got.unvignetted, want.unvignetted)
    # the same through the bare C ABI without an input mask (one bulk copy fewer per tile)
    import torch

    rin, rout = _lib.RaysIn(), _lib.RaysOut()
    rin.n_axes = 1
    rin.dims[0] = n
    outs = {name: torch.empty(n, dtype=torch.float64, device=cuda_device) for name in _lib.FIELDS}
    mask = torch.empty(n, dtype=torch.uint8, device=cuda_device)
    for f, name in enumerate(_lib.FIELDS):
        rin.field[f] = first.fields[name].data_ptr()
        rin.stride[f][0] = 1
        rout.field[f] = outs[name].data_ptr()
    rin.unvignetted = None
    rout.unvignetted = mask.data_ptr()
    _lib.check(
        _lib.lib().optk_trace(
            compiled.handle, 0, C.byref(rin), C.byref(rout), 0, compiled.n_surface, 1, 0, 0, None, None, None, None
        )
    )
    torch.cuda.synchronize()
    for name in want.fields:
        assert torch_equal(outs[name], want.fields[name]), name
    assert torch_equal(mask, want.unvignetted)
    # and against the oracle
    r0, _ = configs.flatten_rays(rays)
    state = ora.propagate_rays(system.surfaces_all, {k: v.reshape(-1) for k, v in r0.items()}, extended=True)
    host = got.to_host()
    axes = tuple(rays.shape)
    flat = lambda a: na.as_named_array(a).numpy(axes).reshape(-1)  # noqa: E731
    dev = dict(
        wavelength=flat(host.wavelength), px=flat(host.position.x), py=flat(host.position.y), pz=flat(host.position.z),
        dx=flat(host.direction.x), dy=flat(host.direction.y), dz=flat(host.direction.z), intensity=flat(host.intensity),
        attenuation=flat(host.attenuation), index_refraction=flat(host.index_refraction),
        unvignetted=flat(host.unvignetted).astype(bool),
    )
    parity.compare_states(dev, state)


def torch_equal(a, b) -> bool:
    import torch

    a, b = a.reshape(-1), b.reshape(-1)
    if a.dtype.is_floating_point:
        return bool(torch.equal(a.view(torch.int64), b.view(torch.int64)))  # NaNs compare by bit pattern
    return bool(torch.equal(a, b))
