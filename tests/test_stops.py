"""
The stop solver (``optika_b200/_stops.py``; reference
``optika/systems/_sequential.py:396-678, 748-789``) driven by the NumPy oracle on
CPU and by the device engine on GPU.  Pins: the reference's own
``test_field_max_matches_source_aperture`` (``optika/systems/_sequential_test.py:538-597``)
and the geometry of the Newtonian doc example.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import transformations as tf
from optika_b200 import _stops

import configs
from oracle_backend import OracleBackend

RADIUS_FIELD = 0.05 * u.deg


def reference_newtonian_test_system():
    """``_system_newtonian`` of ``optika/systems/_sequential_test.py:538-578``."""
    grid = optika.vectors.ObjectVectorArray(
        wavelength=500 * u.nm,
        field=na.Cartesian2dVectorLinearSpace(-1, 1, axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=5, centers=True),
        pupil=na.Cartesian2dVectorLinearSpace(-1, 1, axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=5, centers=True),
    )
    return optika.systems.SequentialSystem(
        object=optika.surfaces.Surface(
            name="source",
            aperture=optika.apertures.CircularAperture(radius=np.sin(RADIUS_FIELD), angular=True),
            is_field_stop=True,
        ),
        surfaces=[
            optika.surfaces.Surface(
                name="primary",
                sag=optika.sags.SphericalSag(radius=-2000 * u.mm),
                material=optika.materials.Mirror(),
                aperture=optika.apertures.CircularAperture(radius=50 * u.mm),
                transformation=tf.Cartesian3dTranslation(z=500 * u.mm),
            ),
            optika.surfaces.Surface(
                name="aperture",
                aperture=optika.apertures.CircularAperture(radius=10 * u.mm),
                transformation=tf.Cartesian3dTranslation(z=250 * u.mm),
                is_pupil_stop=True,
            ),
        ],
        sensor=optika.sensors.ImagingSensor(
            name="sensor",
            width_pixel=15 * u.um,
            axis_pixel=na.Cartesian2dVectorArray("detector_x", "detector_y"),
            num_pixel=na.Cartesian2dVectorArray(128, 128),
            transformation=tf.Cartesian3dTranslation(z=-500 * u.mm),
        ),
        grid_input=grid,
    )


def test_anchor_surface():
    # optika/systems/_sequential_test.py:507-532
    S = optika.surfaces.Surface
    first, last = S(name="first"), S(name="last")
    mirror = S(name="mirror", material=optika.materials.Mirror())
    curved = S(name="curved", sag=optika.sags.SphericalSag(radius=-100 * u.mm))
    grating = S(name="grating", rulings=optika.rulings.Rulings(spacing=1 * u.um, diffraction_order=1))
    flat = S(name="flat")
    assert _stops._anchor_surface([first, flat, mirror, last]) is mirror
    assert _stops._anchor_surface([first, curved, last]) is curved
    assert _stops._anchor_surface([first, grating, last]) is grating
    assert _stops._anchor_surface([first, flat, last]) is last


def check_field_max(backend):
    system = reference_newtonian_test_system()
    result = system.field_max(backend=backend)
    # _sequential_test.py:591-597
    assert abs(float(result.x) - RADIUS_FIELD) < 1e-6 * u.deg
    assert abs(float(result.y) - RADIUS_FIELD) < 1e-6 * u.deg
    lo = system.field_min(backend=backend)
    assert abs(float(lo.x) + RADIUS_FIELD) < 1e-6 * u.deg
    return system


def test_field_max_matches_source_aperture_oracle_backend():
    system = check_field_max(OracleBackend)
    # normalised [-1, 1] grids map onto the solved extents (_sequential.py:748-789)
    grid = system.denormalize(system.grid_input, backend=OracleBackend)
    assert np.allclose(grid.field.x.ndarray / u.deg, [-0.04, -0.02, 0.0, 0.02, 0.04], atol=1e-6)
    pupil_max = system.pupil_max(backend=OracleBackend)
    assert np.allclose(grid.pupil.x.ndarray, np.array([-0.8, -0.4, 0, 0.4, 0.8]) * float(pupil_max.x), atol=1e-9)
    # the entrance pupil is the image of the 10 mm stop through the R = -2000 mirror
    assert 10.0 < float(pupil_max.x) < 20.0


def test_newtonian_doc_example_field_of_view_oracle_backend():
    """Sensor (field stop) half width 128 * 20 um / 2 = 1.28 mm at f = 200 mm -> 0.3667 deg."""
    system = configs.newtonian(num_field=3, num_pupil=3)
    fm = system.field_max(backend=OracleBackend)
    expected = np.arctan(1.28 / 200.0)
    # the edge rays of an f/2.5 paraboloid carry coma, so the extreme field angle is within a
    # percent of (not equal to) the paraxial value
    assert abs(float(fm.x) - expected) < 0.01 * expected
    assert abs(float(fm.y) - expected) < 0.01 * expected
    pm = system.pupil_max(backend=OracleBackend)
    # the primary (+-40 mm) is the pupil stop; at the object plane, 200 mm in front of it, the
    # beams of the extreme fields have walked by 200 mm * tan(field_max) = 1.28 mm
    assert 40.0 < float(pm.x) < 41.3 and 40.0 < float(pm.y) < 41.3


def test_missing_pupil_stop_raises():
    system = configs.newtonian(num_field=2, num_pupil=2)
    for s in system.surfaces:
        s.is_pupil_stop = False
    with pytest.raises(ValueError):
        system.field_max(backend=OracleBackend)


@pytest.mark.gpu
def test_device_backend_matches_oracle_backend(cuda_device):
    system = check_field_max(None)  # the device engine
    _, dev = system.rayfunction_stops(samples_pupil_stop=21, samples_field_stop=21)
    _, ora_rays = system.rayfunction_stops(samples_pupil_stop=21, samples_field_stop=21, backend=OracleBackend)
    for a, b in ((dev.position.x, ora_rays.position.x), (dev.direction.y, ora_rays.direction.y)):
        assert np.allclose(a.numpy(tuple(b.shape)), b.ndarray, rtol=0, atol=1e-8)


@pytest.mark.gpu
def test_normalized_raytrace_on_device(cuda_device):
    """raytrace with normalised field / pupil coordinates (the reference's default) end to end."""
    system = reference_newtonian_test_system()
    result = system.raytrace(normalized_field=True, normalized_pupil=True, accumulate=False)
    rays = result.outputs
    assert rays.shape == {"field_x": 5, "field_y": 5, "pupil_x": 5, "pupil_y": 5}
    # every ray of the normalised grid passes the pupil stop and lands on the sensor
    assert rays.unvignetted.ndarray.mean() > 0.7
    assert np.isfinite(rays.position.x.ndarray).all()


class HostNewtonDeviceBackend(_stops.DeviceBackend):
    """Device traces, Newton iteration on the host (what the device backend did before optk_solve_stops)."""

    solve = None


def _stop_rays(system, backend, samples=11):
    _, rays = system.rayfunction_stops(samples_pupil_stop=samples, samples_field_stop=samples, backend=backend)
    return rays


@pytest.mark.gpu
@pytest.mark.parametrize(
    "make",
    [
        reference_newtonian_test_system,  # unknown = position on the (angular) object surface
        lambda: configs.newtonian(num_field=3, num_pupil=3),  # unknown = direction at the primary, fold mirror between
        lambda: configs.toroidal_vls(num_field=3, num_pupil=3, num_wavelength=2),  # toroid + variable-spacing grating
        lambda: configs.misaligned_telescope(num_field=3, num_pupil=3, num_pixel=64, num_tilt=3),  # configuration axis
    ],
    ids=["reference_test_system", "newtonian", "toroidal_vls", "misaligned_telescope"],
)
def test_device_newton_matches_host_newton(cuda_device, make):
    """
    ``optk_solve_stops`` (one thread per unknown ray) against the host
    iteration of ``_stops._newton`` with the same device traces: both stop at residuals below
    1e-9 x scale, so the solved rays agree far inside the 1e-6 deg / 1e-6 mm the reference's own
    test asks for (``_sequential_test.py:591-597``).
    """
    system = make()
    device_rays = _stop_rays(system, None)
    host_rays = _stop_rays(system, HostNewtonDeviceBackend)
    assert device_rays.shape == host_rays.shape
    for name in ("position", "direction"):
        for c in "xyz":
            a = getattr(getattr(device_rays, name), c)
            b = getattr(getattr(host_rays, name), c)
            b = na.broadcast_to(na.as_named_array(b), device_rays.shape).ndarray
            a = na.broadcast_to(na.as_named_array(a), device_rays.shape).ndarray
            scale = max(1.0, float(np.abs(b).max()))
            assert np.isfinite(a).all()
            assert np.abs(a - b).max() <= 1e-8 * scale, (name, c)


@pytest.mark.gpu
def test_device_newton_residuals_are_below_the_tolerance(cuda_device):
    """Trace the solved rays independently: they arrive on the target grid within max_abs_error."""
    from optika_b200 import _engine, propagators
    from optika_b200.rays import RayVectorArray

    system = configs.newtonian(num_field=3, num_pupil=3)
    surfaces = system.surfaces_all
    first = [i for i, s in enumerate(surfaces) if s.is_pupil_stop][0]
    last = [i for i, s in enumerate(surfaces) if s.is_field_stop][0]
    assert first < last
    subsystem = surfaces[first : last + 1]
    px = na.linspace(-30.0, 30.0, axis="a", num=7)
    py = na.linspace(-20.0, 25.0, axis="b", num=5)
    z = subsystem[0].sag(na.Cartesian3dVectorArray(px, py, 0.0 * (px + py)))
    rays = RayVectorArray(
        wavelength=500 * u.nm,
        position=na.Cartesian3dVectorArray(px, py, z),
        direction=na.Cartesian3dVectorArray(0.0, 0.0, 1.0),
    )
    rays = subsystem[0].transformation(rays)
    target = na.Cartesian2dVectorArray(na.linspace(-1.0, 1.0, axis="t", num=3), na.ScalarArray(np.array(0.25), ()))
    shape_ = na.shape_broadcasted(rays.position, target)
    x0 = na.broadcast_to(na.as_named_array(0.0), shape_)
    solved = _engine.solve_stops(subsystem, rays, "direction", x0, x0, target, "position", 1e-6, 1e-9)
    assert solved is not None
    trial = dataclasses_replace(rays, direction=na.Cartesian3dVectorArray(*solved))
    out = propagators.propagate_rays(subsystem[1:], trial)
    out = subsystem[-1].transformation.inverse(out)
    ex = na.broadcast_to(out.position.x - target.x, shape_).ndarray
    ey = na.broadcast_to(out.position.y - target.y, shape_).ndarray
    assert np.abs(ex).max() <= 2e-9 and np.abs(ey).max() <= 2e-9
    # a target that no ray reaches within one iteration is reported the way the reference does
    with pytest.raises(ValueError, match="Max iterations"):
        _engine.solve_stops(subsystem, rays, "direction", x0, x0, target, "position", 1e-6, 1e-9, max_iterations=1)


def dataclasses_replace(obj, **kwargs):
    import dataclasses

    return dataclasses.replace(obj, **kwargs)
