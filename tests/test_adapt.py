"""
Real optika objects without optika: the adapter (``optika_b200/adapt.py``) duck-types on the public
attributes the reference's objects have (``.unit`` / ``.to_value`` of astropy quantities, ``.ndarray`` /
``.axes`` of named arrays, dataclass field names, class names).  Minimal stand-ins with ONLY those
attributes rebuild the Newtonian telescope of the reference's docstring
(``optika/systems/_sequential.py:1853-1951``) in mixed units; it must lower to the same
``optk_surface_t`` bytes as the same system written with the native classes.
"""

import dataclasses

import numpy as np
import pytest

import configs
from optika_b200 import _lowering, adapt, named as na


class Unit:
    SCALE = {"mm": ("length", 1.0), "m": ("length", 1e3), "cm": ("length", 10.0), "um": ("length", 1e-3),
             "nm": ("length", 1e-6), "rad": ("angle", 1.0), "deg": ("angle", np.pi / 180), "": ("dimensionless", 1.0),
             "s": ("time", 1.0)}

    __array_ufunc__ = None  # ndarray * unit defers to __rmul__, as with astropy units

    def __init__(self, name):
        self.name = name
        self.physical_type, self.scale = self.SCALE[name]

    def __rmul__(self, value):
        return Quantity(value, self)


class Quantity:
    """What the adapter may rely on: ``.unit`` and ``.to_value(unit)`` (raising for another physical type)."""

    def __init__(self, value, unit):
        self._value, self.unit = np.asarray(value, dtype=float), unit

    def to_value(self, unit):
        kind, scale = Unit.SCALE.get(unit, (None, None))
        if kind != self.unit.physical_type:
            raise ValueError(f"'{self.unit.name}' and '{unit}' are not convertible")  # astropy: UnitConversionError
        return self._value * (self.unit.scale / scale)

    def __neg__(self):
        return Quantity(-self._value, self.unit)


class ScalarArray:
    """``named_arrays.ScalarArray``: ``.ndarray`` (here a quantity) and ``.axes``."""

    def __init__(self, ndarray, axes):
        self.ndarray, self.axes = ndarray, (axes,) if isinstance(axes, str) else tuple(axes)


@dataclasses.dataclass
class Cartesian2dVectorArray:
    x: object = 0
    y: object = 0


@dataclasses.dataclass
class Cartesian3dVectorArray:
    x: object = 0
    y: object = 0
    z: object = 0


@dataclasses.dataclass
class Translation:
    vector: Cartesian3dVectorArray


@dataclasses.dataclass
class Cartesian3dRotationY:
    angle: object


@dataclasses.dataclass
class Cartesian3dRotationX:
    angle: object


@dataclasses.dataclass
class TransformationList:
    transformations: list


# stand-ins for the reference's element classes: same names, same field names, nothing else
@dataclasses.dataclass
class ParabolicSag:
    focal_length: object
    transformation: object = None


@dataclasses.dataclass
class RectangularAperture:
    half_width: object
    active: bool = True
    inverted: bool = False
    transformation: object = None


@dataclasses.dataclass
class Mirror:
    pass


@dataclasses.dataclass
class Surface:
    name: str = None
    sag: object = None
    material: object = None
    aperture: object = None
    rulings: object = None
    is_field_stop: bool = False
    is_pupil_stop: bool = False
    transformation: object = None


@dataclasses.dataclass
class ImagingSensor:
    name: str = None
    width_pixel: object = None
    axis_pixel: object = None
    num_pixel: object = None
    timedelta_exposure: object = None
    transformation: object = None
    is_field_stop: bool = False


@dataclasses.dataclass
class SequentialSystem:
    surfaces: list = None
    sensor: object = None
    object: object = None


mm, m, cm, um, deg, rad = Unit("mm"), Unit("m"), Unit("cm"), Unit("um"), Unit("deg"), Unit("rad")


def reference_newtonian(tilt=None):
    """The doc example, deliberately in mixed units (metres, centimetres, degrees)."""
    fold_transformation = TransformationList([
        Cartesian3dRotationY(135 * deg), Translation(Cartesian3dVectorArray(0 * mm, 0 * mm, 0.05 * m)),
    ])
    primary_transformation = Translation(Cartesian3dVectorArray(z=20 * cm))
    if tilt is not None:
        primary_transformation = TransformationList([Cartesian3dRotationX(tilt), primary_transformation])
    return SequentialSystem(
        surfaces=[
            Surface(name="front"),
            Surface(name="obscuration", aperture=RectangularAperture(2.5 * cm, inverted=True), transformation=fold_transformation),
            Surface(name="mirror", sag=ParabolicSag(focal_length=-(0.2 * m)), aperture=RectangularAperture(40 * mm),
                    material=Mirror(), is_pupil_stop=True, transformation=primary_transformation),
            Surface(name="fold_mirror", aperture=RectangularAperture(25 * mm), material=Mirror(), transformation=fold_transformation),
        ],
        sensor=ImagingSensor(
            name="sensor", width_pixel=20 * um, axis_pixel=Cartesian2dVectorArray("detector_x", "detector_y"),
            num_pixel=Cartesian2dVectorArray(128, 128), timedelta_exposure=1 * Unit("s"), is_field_stop=True,
            transformation=TransformationList([
                Cartesian3dRotationY(-90 * deg), Translation(Cartesian3dVectorArray(x=-(50 * mm), z=5 * cm)),
            ]),
        ),
    )


def table_bytes(system):
    table, shape_ = _lowering.lower_system(system.surfaces_all)
    return _lowering.table_key(table), shape_


def test_adapted_newtonian_lowers_to_the_same_bytes_as_the_native_one():
    adapted = adapt.from_reference(reference_newtonian())
    native = configs.newtonian()
    assert type(adapted).__module__.startswith("optika_b200")
    (got, got_shape), (want, want_shape) = table_bytes(adapted), table_bytes(native)
    assert got_shape == want_shape == {}
    assert len(got) == len(want)
    table_got, _ = _lowering.lower_system(adapted.surfaces_all)
    table_want, _ = _lowering.lower_system(native.surfaces_all)
    for k in range(len(table_want)):
        for name in ("sag", "aperture"):
            assert np.allclose(list(getattr(table_got[k], name)), list(getattr(table_want[k], name)), rtol=1e-15, atol=0)
        assert np.allclose(list(table_got[k].transform.r), list(table_want[k].transform.r), rtol=0, atol=1e-16)
        assert np.allclose(list(table_got[k].transform.t), list(table_want[k].transform.t), rtol=1e-15, atol=1e-13)
        for name in ("sag_kind", "material_kind", "aperture_kind", "ruling_kind", "flags"):
            assert getattr(table_got[k], name) == getattr(table_want[k], name), (k, name)
    assert adapted.surfaces[2].is_pupil_stop and adapted.sensor.is_field_stop


def test_named_quantity_arrays_become_configuration_axes():
    tilt = ScalarArray(np.linspace(-30, 30, 4) / 3600 * deg, "misalign")  # a named array of a quantity
    adapted = adapt.from_reference(reference_newtonian(tilt=tilt))
    native = configs.misaligned_telescope(num_tilt=4)
    table_got, shape_got = _lowering.lower_system(adapted.surfaces_all)
    table_want, shape_want = _lowering.lower_system(native.surfaces_all)
    assert shape_got == shape_want == {"misalign": 4}
    for k in range(len(table_want)):
        assert np.allclose(list(table_got[k].transform.r), list(table_want[k].transform.r), rtol=0, atol=1e-15)
        assert np.allclose(list(table_got[k].transform.t), list(table_want[k].transform.t), rtol=1e-15, atol=1e-13)


def test_engine_units_follow_the_quantitys_own_physical_type():
    assert adapt.engine_value(2 * m) == 2000.0
    assert adapt.engine_value(90 * deg) == pytest.approx(np.pi / 2)
    assert adapt.engine_value(Quantity(0.25, Unit(""))) == 0.25
    named = adapt.from_reference(ScalarArray(np.array([1.0, 2.0]) * cm, "radius"))
    assert isinstance(named, na.ScalarArray) and named.axes == ("radius",) and np.array_equal(named.ndarray, [10.0, 20.0])
    with pytest.raises(NotImplementedError, match="no counterpart"):
        adapt.from_reference(object())
