"""
Adapter for REAL optika objects (the drop-in requirement of BASELINE ``north_star``).

``optika_b200`` mirrors the reference's classes, but a user of the reference holds objects made of
``optika.surfaces.Surface`` / ``optika.sags.*`` / ... whose parameters are ``astropy.units.Quantity``
and ``named_arrays`` arrays (``optika/surfaces.py:282-395``).  :func:`from_reference` converts such an
object graph into the native classes by DUCK TYPING -- nothing here imports ``optika``, ``astropy`` or
``named_arrays`` (none of them is installable where this engine is built):

* a class is recognised by its NAME (``type(obj).__name__``: ``ParabolicSag``, ``RectangularAperture``,
  ``Rulings``, ``ImagingSensor``, ``SequentialSystem`` ...), and every dataclass field the native class
  shares with it is converted recursively;
* a quantity is anything with ``.unit`` and ``.to_value(unit)``: lengths become millimetres, angles
  radians, reciprocal lengths 1 / mm, areas mm^2, dimensionless values plain floats -- decided by which
  conversion the object itself accepts (astropy raises ``UnitConversionError`` for the others);
* a named array is anything with ``.ndarray`` and ``.axes``; a vector anything with ``.x``, ``.y``
  (``.z``) that is not a surface;
* transformations are recognised by name: ``Translation`` / ``Cartesian3dTranslation`` (``.vector`` or
  ``.x .y .z``), ``Cartesian3dRotationX/Y/Z`` (``.angle``), ``TransformationList`` (``.transformations``).

The result lowers to exactly the ``optk_surface_t`` bytes the same system written with the native
classes lowers to (``tests/test_adapt.py``).
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na

__all__ = ["from_reference", "engine_value"]

_UNITS = ("mm", "rad", "", "1 / mm", "mm2", "s", "electron", "photon")  # tried in this order


def _native_classes() -> dict:
    from . import apertures, materials, rays, rulings, sags, sensors, surfaces, systems, transformations, vectors
    from .materials import _layers, _materials, profiles

    found = {}
    for module in (apertures, materials, _layers, _materials, profiles, rays, rulings, sags, sensors, surfaces, systems,
                   transformations, vectors):
        for name, cls in vars(module).items():
            if isinstance(cls, type) and dataclasses.is_dataclass(cls) and not name.startswith("_"):
                found.setdefault(name, cls)
    return found


_CLASSES = None


def _is_quantity(a) -> bool:
    return hasattr(a, "unit") and hasattr(a, "to_value")


def _is_named(a) -> bool:
    return hasattr(a, "ndarray") and hasattr(a, "axes") and not dataclasses.is_dataclass(a)


def engine_value(a):
    """A quantity in engine units (mm, rad, 1 / mm, mm^2, plain numbers), as a float or an ndarray."""
    last = None
    for unit in _UNITS:
        try:
            v = a.to_value(unit)
        except Exception as e:  # astropy: UnitConversionError; the wrong physical type for this unit
            last = e
            continue
        v = np.asarray(v)
        return v.item() if v.ndim == 0 else v
    raise ValueError(f"cannot express {a!r} in engine units (mm, rad, 1 / mm, mm^2, dimensionless)") from last


def _vector(obj):
    comps = [from_reference(getattr(obj, c)) for c in ("x", "y", "z") if hasattr(obj, c)]
    return na.Cartesian3dVectorArray(*comps) if len(comps) == 3 else na.Cartesian2dVectorArray(*comps)


def _transformation(obj):
    from . import transformations as tf

    name = type(obj).__name__
    if name == "TransformationList":
        return tf.TransformationList([from_reference(t) for t in obj.transformations])
    if name in ("Cartesian3dRotationX", "Cartesian3dRotationY", "Cartesian3dRotationZ"):
        return getattr(tf, name)(from_reference(obj.angle))
    if name in ("Translation", "Cartesian3dTranslation"):
        v = obj.vector if hasattr(obj, "vector") else obj
        get = lambda c: from_reference(getattr(v, c)) if hasattr(v, c) else 0  # noqa: E731
        return tf.Cartesian3dTranslation(x=get("x"), y=get("y"), z=get("z"))
    if name == "IdentityTransformation":
        return tf.IdentityTransformation()
    raise NotImplementedError(f"transformation {name} is not supported by the device engine")


def from_reference(obj):
    """The native (``optika_b200``) counterpart of a reference object, recursively."""
    global _CLASSES
    if _CLASSES is None:
        _CLASSES = _native_classes()
    if obj is None or isinstance(obj, (bool, int, float, complex, str, np.generic)):
        return obj
    if type(obj).__module__.startswith("optika_b200"):
        return obj  # already native
    if _is_named(obj):
        nd = obj.ndarray
        nd = engine_value(nd) if _is_quantity(nd) else np.asarray(nd)
        return na.ScalarArray(nd, tuple(obj.axes))
    if _is_quantity(obj):
        return engine_value(obj)
    if isinstance(obj, np.ndarray):
        return obj
    if isinstance(obj, (list, tuple)):
        return type(obj)(from_reference(v) for v in obj)
    if isinstance(obj, dict):
        return {k: from_reference(v) for k, v in obj.items()}
    name = type(obj).__name__
    if "Transformation" in name or "Translation" in name or "Rotation" in name:
        return _transformation(obj)
    cls = _CLASSES.get(name)
    if cls is not None:
        kwargs = {}
        for f in dataclasses.fields(cls):
            if f.init and hasattr(obj, f.name):
                kwargs[f.name] = from_reference(getattr(obj, f.name))
        return cls(**kwargs)
    if hasattr(obj, "x") and hasattr(obj, "y"):
        return _vector(obj)
    raise NotImplementedError(f"{name} has no counterpart in the device engine")
