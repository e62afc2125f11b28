"""
The BASELINE.json configurations, built through the optika-compatible API.
Shared by the parity tests, ``__graft_entry__.smoke()`` and ``bench.py``.
(Not a test module.)
"""

from __future__ import annotations
import numpy as np
import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import transformations as tf

# The grids of these configurations are PHYSICAL (degrees, millimetres).  Floats carry no unit here,
# so -- unlike the reference, which tells normalised from physical coordinates by their unit -- a call
# must say so: ``system.raytrace(**PHYSICAL)``.  The defaults are the reference's (normalised).
PHYSICAL = dict(normalized_field=False, normalized_pupil=False)


def newtonian(num_field: int = 10, num_pupil: int = 32, num_pixel: int = 128):
    """
    cfg 1: the Newtonian telescope of the reference's docstring example,
    ``optika/systems/_sequential.py:1853-1951``.
    Physical (not normalised) grid: field +-0.1 deg, pupil +-40 mm, cell centres.
    """
    primary_mirror_z = 200 * u.mm
    fold_mirror_z = 50 * u.mm
    sensor_x = 50 * u.mm
    front = optika.surfaces.Surface(name="front")
    primary_mirror = optika.surfaces.Surface(
        name="mirror",
        sag=optika.sags.ParabolicSag(focal_length=-(primary_mirror_z - fold_mirror_z + sensor_x)),
        aperture=optika.apertures.RectangularAperture(40 * u.mm),
        material=optika.materials.Mirror(),
        is_pupil_stop=True,
        transformation=tf.Cartesian3dTranslation(z=primary_mirror_z),
    )
    fold_mirror = optika.surfaces.Surface(
        name="fold_mirror",
        aperture=optika.apertures.RectangularAperture(25 * u.mm),
        material=optika.materials.Mirror(),
        transformation=tf.TransformationList(
            [
                tf.Cartesian3dRotationY((90 + 45) * u.deg),
                tf.Cartesian3dTranslation(z=fold_mirror_z),
            ]
        ),
    )
    obscuration = optika.surfaces.Surface(
        name="obscuration",
        aperture=optika.apertures.RectangularAperture(25 * u.mm, inverted=True),
        transformation=fold_mirror.transformation,
    )
    sensor = optika.sensors.ImagingSensor(
        name="sensor",
        width_pixel=20 * u.um,
        axis_pixel=na.Cartesian2dVectorArray("detector_x", "detector_y"),
        num_pixel=na.Cartesian2dVectorArray(num_pixel, num_pixel),
        timedelta_exposure=1 * u.s,
        transformation=tf.TransformationList(
            [
                tf.Cartesian3dRotationY(-90 * u.deg),
                tf.Cartesian3dTranslation(x=-sensor_x, z=fold_mirror_z),
            ]
        ),
        is_field_stop=True,
    )
    field = na.Cartesian2dVectorLinearSpace(
        start=-0.1 * u.deg, stop=0.1 * u.deg,
        axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=num_field, centers=True,
    )
    pupil = na.Cartesian2dVectorLinearSpace(
        start=-40 * u.mm, stop=40 * u.mm,
        axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=num_pupil, centers=True,
    )
    return optika.systems.SequentialSystem(
        surfaces=[front, obscuration, primary_mirror, fold_mirror],
        sensor=sensor,
        grid_input=optika.vectors.ObjectVectorArray(wavelength=500 * u.nm, field=field, pupil=pupil),
    )


def spherical_grating(
    num_field: int = 8, num_pupil: int = 16, num_wavelength: int = 4, num_pixel: int = 2048
):
    """
    cfg 2: object -> concave spherical grating (R = -1000 mm, 1200 lines/mm,
    order 1, circular aperture 50 mm) at z = 1000 mm -> sensor at the m = 1
    focus.  Wavelengths uniform in 17-63 nm.
    """
    radius = 1000 * u.mm
    grating = optika.surfaces.Surface(
        name="grating",
        sag=optika.sags.SphericalSag(radius=-radius),
        material=optika.materials.Mirror(),
        rulings=optika.rulings.Rulings(spacing=(1 / 1200) * u.mm, diffraction_order=1),
        aperture=optika.apertures.CircularAperture(50 * u.mm),
        is_pupil_stop=True,
        transformation=tf.Cartesian3dTranslation(z=radius),
    )
    # Rowland-like geometry: the first-order beam of the central wavelength leaves the
    # grating vertex at sin(beta) = m lambda / d; put the sensor on that chief ray.
    w0 = 40 * u.nm
    sin_beta = 1 * w0 / ((1 / 1200) * u.mm)
    beta = np.arcsin(sin_beta)
    distance = 500 * u.mm
    sensor = optika.sensors.ImagingSensor(
        name="sensor",
        width_pixel=15 * u.um,
        axis_pixel=na.Cartesian2dVectorArray("detector_x", "detector_y"),
        num_pixel=na.Cartesian2dVectorArray(num_pixel, num_pixel),
        transformation=tf.TransformationList(
            [
                tf.Cartesian3dRotationY(np.pi + beta),
                tf.Cartesian3dTranslation(
                    x=-distance * np.sin(beta), z=radius - distance * np.cos(beta)
                ),
            ]
        ),
        is_field_stop=True,
    )
    field = na.Cartesian2dVectorLinearSpace(
        start=-0.05 * u.deg, stop=0.05 * u.deg,
        axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=num_field, centers=True,
    )
    pupil = na.Cartesian2dVectorLinearSpace(
        start=-45 * u.mm, stop=45 * u.mm,
        axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=num_pupil, centers=True,
    )
    wavelength = na.linspace(17 * u.nm, 63 * u.nm, axis="wavelength", num=num_wavelength)
    return optika.systems.SequentialSystem(
        surfaces=[grating],
        sensor=sensor,
        grid_input=optika.vectors.ObjectVectorArray(wavelength=wavelength, field=field, pupil=pupil),
    )


def toroidal_vls(num_field: int = 6, num_pupil: int = 12, num_wavelength: int = 3):
    """
    cfg 3: EUV slitless spectrograph: object -> octagonal field stop ->
    toroidal variable-line-spacing grating (polynomial spacing, powers 0, 1, 2)
    with a rectangular aperture -> sensor 2048 x 1024.
    """
    stop = optika.surfaces.Surface(
        name="field_stop",
        aperture=optika.apertures.OctagonalAperture(radius=20 * u.mm),
        is_field_stop=True,
        transformation=tf.Cartesian3dTranslation(z=100 * u.mm),
    )
    z_grating = 1500 * u.mm
    grating = optika.surfaces.Surface(
        name="grating",
        # the reference's toroid formula needs a positive radius of rotation
        # (optika/sags/_toroidal.py:56-57), so the concave side faces -z of the
        # surface frame and the surface is flipped to face the incoming light
        sag=optika.sags.ToroidalSag(radius=1000 * u.mm, radius_of_rotation=1020 * u.mm),
        material=optika.materials.Mirror(),
        rulings=optika.rulings.Rulings(
            spacing=optika.rulings.Polynomial1dRulingSpacing(
                coefficients={
                    0: (1 / 2400) * u.mm,
                    1: 2e-8,
                    2: 1e-11 / u.mm,
                },
                normal=na.Cartesian3dVectorArray(1, 0, 0),
            ),
            diffraction_order=1,
        ),
        aperture=optika.apertures.RectangularAperture(
            na.Cartesian2dVectorArray(30 * u.mm, 20 * u.mm)
        ),
        is_pupil_stop=True,
        transformation=tf.TransformationList(
            [tf.Cartesian3dRotationY(180 * u.deg), tf.Cartesian3dTranslation(z=z_grating)]
        ),
    )
    w0 = 30 * u.nm
    beta = np.arcsin(w0 / ((1 / 2400) * u.mm))
    distance = 510 * u.mm
    sensor = optika.sensors.ImagingSensor(
        name="sensor",
        width_pixel=15 * u.um,
        axis_pixel=na.Cartesian2dVectorArray("detector_x", "detector_y"),
        num_pixel=na.Cartesian2dVectorArray(2048, 1024),
        transformation=tf.TransformationList(
            [
                tf.Cartesian3dRotationY(np.pi + beta),
                tf.Cartesian3dTranslation(x=-distance * np.sin(beta), z=z_grating - distance * np.cos(beta)),
            ]
        ),
    )
    field = na.Cartesian2dVectorLinearSpace(
        start=-0.2 * u.deg, stop=0.2 * u.deg,
        axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=num_field, centers=True,
    )
    pupil = na.Cartesian2dVectorLinearSpace(
        start=-22 * u.mm, stop=22 * u.mm,
        axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=num_pupil, centers=True,
    )
    wavelength = na.linspace(25 * u.nm, 35 * u.nm, axis="wavelength", num=num_wavelength)
    return optika.systems.SequentialSystem(
        surfaces=[stop, grating],
        sensor=sensor,
        grid_input=optika.vectors.ObjectVectorArray(wavelength=wavelength, field=field, pupil=pupil),
    )


def misaligned_telescope(num_field: int = 6, num_pupil: int = 12, num_pixel: int = 256, num_tilt: int = 4):
    """
    cfg 5: the Newtonian telescope with a named configuration axis ``misalign``:
    `num_tilt` tilts of the primary about x within +-30 arcsec.
    """
    system = newtonian(num_field=num_field, num_pupil=num_pupil, num_pixel=num_pixel)
    tilt = na.linspace(-30 * u.arcsec, 30 * u.arcsec, axis="misalign", num=num_tilt)
    primary = system.surfaces[2]
    primary.transformation = tf.TransformationList(
        [tf.Cartesian3dRotationX(tilt), tf.Cartesian3dTranslation(z=200 * u.mm)]
    )
    return system


def telescope_4k(num_tilt: int = 8, num_pixel: int = 4096):
    """
    cfg 5 at full size: the Newtonian layout of cfg 1 scaled to a 4096 x 4096, 15 um sensor (61.4 mm)
    so that the scene FILLS the detector -- a 320 mm f/10 parabola (f = 3200 mm), field of view
    +-0.55 deg, flat fold 250 mm in front of the focus, the obscuration the fold casts on the way in;
    named configuration axis ``misalign`` = `num_tilt` tilts of the primary about x within +-30 arcsec
    (each shifts the image by up to 62 pixels).  Coma at the corner of the field is ~5 pixels, so the
    pupil samples of one field cell spread over a handful of pixels, as in a real instrument.
    """
    focal = 3200 * u.mm
    sensor_x = 250 * u.mm
    fold_z = 200 * u.mm
    primary_z = fold_z + (focal - sensor_x)
    front = optika.surfaces.Surface(name="front")
    tilt = na.linspace(-30 * u.arcsec, 30 * u.arcsec, axis="misalign", num=num_tilt) if num_tilt > 1 else 0.0
    primary = optika.surfaces.Surface(
        name="mirror",
        sag=optika.sags.ParabolicSag(focal_length=-focal),
        aperture=optika.apertures.RectangularAperture(160 * u.mm),
        material=optika.materials.Mirror(),
        is_pupil_stop=True,
        transformation=tf.TransformationList(
            [tf.Cartesian3dRotationX(tilt), tf.Cartesian3dTranslation(z=primary_z)]
        ),
    )
    fold = optika.surfaces.Surface(
        name="fold_mirror",
        aperture=optika.apertures.RectangularAperture(na.Cartesian2dVectorArray(62 * u.mm, 45 * u.mm)),
        material=optika.materials.Mirror(),
        transformation=tf.TransformationList(
            [tf.Cartesian3dRotationY((90 + 45) * u.deg), tf.Cartesian3dTranslation(z=fold_z)]
        ),
    )
    obscuration = optika.surfaces.Surface(
        name="obscuration",
        aperture=optika.apertures.RectangularAperture(
            na.Cartesian2dVectorArray(62 * u.mm, 45 * u.mm), inverted=True
        ),
        transformation=fold.transformation,
    )
    sensor = optika.sensors.ImagingSensor(
        name="sensor",
        width_pixel=15 * u.um,
        axis_pixel=na.Cartesian2dVectorArray("detector_x", "detector_y"),
        num_pixel=na.Cartesian2dVectorArray(num_pixel, num_pixel),
        timedelta_exposure=1 * u.s,
        transformation=tf.TransformationList(
            [tf.Cartesian3dRotationY(-90 * u.deg), tf.Cartesian3dTranslation(x=-sensor_x, z=fold_z)]
        ),
        is_field_stop=True,
    )
    half = np.arctan(0.5 * num_pixel * 15e-3 / 3200.0)
    field = na.Cartesian2dVectorLinearSpace(
        start=-half, stop=half, axis=na.Cartesian2dVectorArray("field_x", "field_y"), num=8, centers=True,
    )
    pupil = na.Cartesian2dVectorLinearSpace(
        start=-160 * u.mm, stop=160 * u.mm, axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=8, centers=True,
    )
    return optika.systems.SequentialSystem(
        surfaces=[front, obscuration, primary, fold],
        sensor=sensor,
        grid_input=optika.vectors.ObjectVectorArray(wavelength=500 * u.nm, field=field, pupil=pupil),
    )


def flatten_rays(rays) -> dict:
    """Host RayVectorArray -> the oracle's dict of equally shaped arrays (C order of rays.shape)."""
    shape_ = rays.shape
    dims = tuple(shape_.values())

    def get(v, dtype=float):
        return np.broadcast_to(na.aligned(na.as_named_array(v), shape_), dims).astype(dtype)

    return dict(
        wavelength=get(rays.wavelength),
        px=get(rays.position.x), py=get(rays.position.y), pz=get(rays.position.z),
        dx=get(rays.direction.x), dy=get(rays.direction.y), dz=get(rays.direction.z),
        intensity=get(rays.intensity),
        attenuation=get(rays.attenuation),
        index_refraction=get(rays.index_refraction),
        unvignetted=get(rays.unvignetted, bool),
    ), shape_
