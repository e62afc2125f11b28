"""
Lowering pass: optical elements -> the device surface table.

Turns a list of :class:`optika_b200.surfaces.Surface`-like objects (the
``SequentialSystem.surfaces_all`` list of ``optika/systems/_sequential.py:93-108``)
into ``[n_config][n_surface]`` packed ``optk_surface_t`` records
(``include/optk.h``).  Named configuration axes carried by any parameter
(``optika/surfaces.py:386-395``, ``optika/systems/_sequential.py:2125-2132``)
are broadcast into the configuration shape and evaluated one configuration at a
time; transformations are composed into one affine per element.
"""

from __future__ import annotations
import ctypes as C
import numpy as np
from . import named as na
from . import units as u
from . import _lib as L

__all__ = ["config_shape", "lower_surface", "lower_system", "affine_struct"]


def config_shape(surfaces) -> dict[str, int]:
    """Broadcast of the named shapes of all surfaces (``_sequential.py:2125-2132``)."""
    return na.broadcast_shapes(*[na.shape(s) for s in surfaces])


def _scalar(value, shape_: dict[str, int], index: tuple) -> float:
    """Value of a (possibly named-array) parameter at one configuration index."""
    if isinstance(value, na.ScalarArray):
        nd = np.broadcast_to(na.aligned(value, shape_), tuple(shape_.values()))
        return nd[index].item() if index else nd.item()
    return value


def affine_struct(transformation, shape_: dict[str, int], index: tuple) -> L.Affine:
    a = L.Affine()
    if transformation is None:
        a.r[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
        a.t[:] = [0, 0, 0]
        return a
    r, t = transformation.affine.numpy(shape_)
    r = r[index] if index else r
    t = t[index] if index else t
    a.r[:] = list(np.asarray(r, dtype=float).reshape(9))
    a.t[:] = list(np.asarray(t, dtype=float).reshape(3))
    return a


def _vector3(v, shape_, index) -> list[float]:
    return [float(_scalar(u.length(c), shape_, index)) for c in (v.x, v.y, v.z)]


def _lower_sag(sag, S: L.Surface, shape_, index):
    name = type(sag).__name__
    get = lambda v: float(_scalar(u.length(v), shape_, index))  # noqa: E731
    if name == "NoSag":
        S.sag_kind = L.SAG_FLAT
    elif name == "SphericalSag":
        S.sag_kind = L.SAG_SPHERICAL
        S.sag[0] = get(sag.radius)
    elif name == "CylindricalSag":
        S.sag_kind = L.SAG_CYLINDRICAL
        S.sag[0] = get(sag.radius)
    elif name == "ConicSag":
        S.sag_kind = L.SAG_CONIC
        S.sag[0] = get(sag.radius)
        S.sag[1] = get(sag.conic)
    elif name == "ParabolicSag":
        S.sag_kind = L.SAG_PARABOLIC
        S.sag[0] = get(sag.focal_length)
    elif name == "ToroidalSag":
        S.sag_kind = L.SAG_TOROIDAL
        S.sag[0] = get(sag.radius)
        S.sag[2] = get(sag.radius_of_rotation)
    else:
        raise NotImplementedError(f"sag {name} is not supported by the device engine")
    t = getattr(sag, "transformation", None)
    if t is not None:
        S.flags |= L.F_SAG_TRANSFORM
    S.sag_transform = affine_struct(t, shape_, index)


def _measured_table(measured, shape_, index, keep: list):
    """
    ``(n, x pointer, y pointer)`` of a measured efficiency (``optika/materials/_materials.py:283-305``,
    ``optika/rulings/_rulings.py:291-313``) at one configuration index; the host arrays are
    appended to `keep` (``optk_system_create`` copies them to the device).
    """
    import ctypes as C

    inputs = measured.inputs
    wavelength = na.as_named_array(u.length(inputs.wavelength))
    if na.as_named_array(getattr(inputs.direction, "x", inputs.direction)).size != 1 or any(
        n != 1 for n in na.shape(inputs.direction).values()
    ):
        raise ValueError("Interpolating over different incidence angles is not supported.")
    if wavelength.ndim != 1:
        raise ValueError(f"wavelength must be one dimensional, got shape {wavelength.shape}")
    (axis,) = wavelength.axes
    full = dict(shape_)
    full[axis] = wavelength.shape[axis]
    values = np.broadcast_to(na.aligned(na.as_named_array(measured.outputs), full), tuple(full.values()))
    x = np.ascontiguousarray(wavelength.ndarray, dtype=np.float64)
    y = np.ascontiguousarray(values[index], dtype=np.float64)
    if np.any(np.diff(x) < 0):  # numpy.interp wants ascending abscissae
        order = np.argsort(x)
        x, y = np.ascontiguousarray(x[order]), np.ascontiguousarray(y[order])
    keep += [x, y]
    return len(x), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)


def _lower_material(material, S: L.Surface, shape_, index, keep: list):
    name = type(material).__name__
    if name in ("Vacuum", "IdealSensorMaterial"):
        S.material_kind = L.MAT_VACUUM
    elif name == "Mirror":
        S.material_kind = L.MAT_MIRROR
    elif name == "MeasuredMirror":
        S.material_kind = L.MAT_MIRROR
        S.material_efficiency = L.EFF_LUT
        S.material_lut_n, S.material_lut_x, S.material_lut_y = _measured_table(
            material.efficiency_measured, shape_, index, keep
        )
    elif name == "MultilayerMirror":
        S.material_kind = L.MAT_MIRROR  # the efficiency is applied by the engine's chained trace
    elif name == "MultilayerFilm":
        S.material_kind = L.MAT_PASS
    elif name == "Glass":
        S.material_kind = L.MAT_GLASS
        for k, v in enumerate((material.b1, material.b2, material.b3, material.c1, material.c2, material.c3)):
            S.material[k] = float(_scalar(v, shape_, index))
    elif name == "_FixedIndex":
        S.material_kind = L.MAT_INDEX_MIRROR if material.is_mirror else L.MAT_INDEX
        S.material[0] = float(_scalar(material.index, shape_, index))
    else:
        raise NotImplementedError(f"material {name} is not supported by the device engine")


_PROFILES = {
    "Rulings": L.PROFILE_IDEAL,
    "SinusoidalRulings": L.PROFILE_SINUSOIDAL,
    "SquareRulings": L.PROFILE_SQUARE,
    "SawtoothRulings": L.PROFILE_SAWTOOTH,
    "TriangularRulings": L.PROFILE_TRIANGULAR,
    "RectangularRulings": L.PROFILE_RECTANGULAR,
    "MeasuredRulings": L.PROFILE_MEASURED,
}


def _lower_rulings(rulings, S: L.Surface, shape_, index, keep: list):
    if rulings is None:
        S.ruling_kind = L.RULING_NONE
        return
    kind = type(rulings).__name__
    if kind not in _PROFILES:
        raise NotImplementedError(f"rulings {kind} are not supported by the device engine")
    S.ruling_profile = _PROFILES[kind]
    if hasattr(rulings, "depth"):
        S.ruling_depth = float(_scalar(u.length(rulings.depth), shape_, index))
    if hasattr(rulings, "ratio_duty"):
        S.ruling_duty = float(_scalar(rulings.ratio_duty, shape_, index))
    if kind == "MeasuredRulings":
        S.ruling_lut_n, S.ruling_lut_x, S.ruling_lut_y = _measured_table(
            rulings.efficiency_measured, shape_, index, keep
        )
    spacing = rulings.spacing_
    name = type(spacing).__name__
    S.ruling_order = float(_scalar(rulings.diffraction_order, shape_, index))
    if name == "ConstantRulingSpacing":
        S.ruling_kind = L.RULING_CONSTANT
        S.ruling_coeff[0] = float(_scalar(u.length(spacing.constant), shape_, index))
        S.ruling_normal[:] = _vector3(spacing.normal, shape_, index)
    elif name == "Polynomial1dRulingSpacing":
        S.ruling_kind = L.RULING_POLYNOMIAL
        S.ruling_normal[:] = _vector3(spacing.normal, shape_, index)
        if len(spacing.coefficients) > L.MAX_COEFF:
            raise ValueError(f"at most {L.MAX_COEFF} polynomial coefficients are supported")
        for k, (power, c) in enumerate(spacing.coefficients.items()):
            if int(power) != power:
                raise ValueError("polynomial powers must be integers")
            S.ruling_power[k] = int(power)
            S.ruling_coeff[k] = float(_scalar(c, shape_, index))
        S.n_coeff = len(spacing.coefficients)
        if spacing.transformation is not None:
            S.flags |= L.F_RULING_TRANSFORM
        S.ruling_transform = affine_struct(spacing.transformation, shape_, index)
    elif name == "HolographicRulingSpacing":
        S.ruling_kind = L.RULING_HOLOGRAPHIC
        S.holo_x1[:] = _vector3(spacing.x1, shape_, index)
        S.holo_x2[:] = _vector3(spacing.x2, shape_, index)
        S.holo_wavelength = float(_scalar(u.length(spacing.wavelength), shape_, index))
        if bool(_scalar(spacing.is_diverging_1, shape_, index)):
            S.flags |= L.F_HOLO_DIVERGING_1
        if bool(_scalar(spacing.is_diverging_2, shape_, index)):
            S.flags |= L.F_HOLO_DIVERGING_2
    else:
        raise NotImplementedError(f"ruling spacing {name} is not supported by the device engine")


def _lower_aperture(aperture, S: L.Surface, shape_, index):
    if aperture is None:
        S.aperture_kind = L.APERTURE_NONE
        return
    name = type(aperture).__name__
    get = lambda v: float(_scalar(u.length(v), shape_, index))  # noqa: E731
    if name == "CircularAperture":
        S.aperture_kind = L.APERTURE_CIRCULAR
        S.aperture[0] = get(aperture.radius)
    elif name == "CircularSectorAperture":
        S.aperture_kind = L.APERTURE_SECTOR
        S.aperture[0] = get(aperture.radius)
        S.aperture[1] = float(_scalar(u.angle(aperture.angle_start), shape_, index))
        S.aperture[2] = float(_scalar(u.angle(aperture.angle_stop), shape_, index))
    elif name == "EllipticalAperture":
        S.aperture_kind = L.APERTURE_ELLIPTICAL
        S.aperture[0] = get(aperture.radius.x)
        S.aperture[1] = get(aperture.radius.y)
    elif name == "RectangularAperture":
        S.aperture_kind = L.APERTURE_RECTANGULAR
        h = aperture.half_width_xy
        S.aperture[0] = get(h.x)
        S.aperture[1] = get(h.y)
    elif hasattr(aperture, "vertices"):
        S.aperture_kind = L.APERTURE_POLYGON
        v = aperture.vertices
        shape_v = dict(vertex=na.shape(v)["vertex"], **shape_)
        vx = np.broadcast_to(na.aligned(na.as_named_array(v.x), shape_v), tuple(shape_v.values()))
        vy = np.broadcast_to(na.aligned(na.as_named_array(v.y), shape_v), tuple(shape_v.values()))
        vx = vx[(slice(None),) + index]
        vy = vy[(slice(None),) + index]
        if len(vx) > L.MAX_VERTICES:
            raise ValueError(f"polygon apertures support at most {L.MAX_VERTICES} vertices")
        S.n_vertices = len(vx)
        for k in range(len(vx)):
            S.vertices_x[k] = float(vx[k])
            S.vertices_y[k] = float(vy[k])
    else:
        raise NotImplementedError(f"aperture {name} is not supported by the device engine")
    if bool(_scalar(aperture.active, shape_, index)):
        S.flags |= L.F_APERTURE_ACTIVE
    if bool(_scalar(aperture.inverted, shape_, index)):
        S.flags |= L.F_APERTURE_INVERTED
    if getattr(aperture, "angular", False):
        S.flags |= L.F_APERTURE_ANGULAR
    if aperture.transformation is not None:
        S.flags |= L.F_APERTURE_TRANSFORM
    S.aperture_transform = affine_struct(aperture.transformation, shape_, index)


def lower_surface(
    surface, shape_: dict[str, int], index: tuple, stages: int = L.STAGE_ALL, keep: list | None = None
) -> L.Surface:
    """
    One surface at one configuration index -> ``optk_surface_t``.  Host arrays the record
    points to (measured-efficiency tables) are appended to `keep`, which must outlive the
    ``optk_system_create`` call.
    """
    S = L.Surface()
    S.stages = stages
    S.flags = 0
    keep = [] if keep is None else keep
    _lower_sag(surface.sag, S, shape_, index)
    _lower_material(surface.material, S, shape_, index, keep)
    _lower_rulings(surface.rulings, S, shape_, index, keep)
    _lower_aperture(surface.aperture, S, shape_, index)
    if surface.transformation is not None:
        S.flags |= L.F_TRANSFORM
    S.transform = affine_struct(surface.transformation, shape_, index)
    return S


def lower_system(surfaces, shape_: dict[str, int] | None = None, stages: int = L.STAGE_ALL):
    """
    ``[n_config][n_surface]`` table as a ctypes array, plus the configuration shape.  The
    host arrays the table points to are returned in ``table.keep``.
    """
    surfaces = list(surfaces)
    if shape_ is None:
        shape_ = config_shape(surfaces)
    n_config = int(np.prod(list(shape_.values()), dtype=np.int64)) if shape_ else 1
    table = _table_type(n_config * len(surfaces))()
    table.keep = []
    # a surface without named axes is the same record in every configuration: lowered once
    # (a tolerance sweep varies one or two surfaces out of many)
    fixed = [None if na.shape(s) else lower_surface(s, {}, (), stages, table.keep) for s in surfaces]
    k = 0
    for index in np.ndindex(*shape_.values()):
        for s, record in zip(surfaces, fixed):
            table[k] = record if record is not None else lower_surface(s, shape_, index, stages, table.keep)
            k += 1
    return table, shape_


_POINTER_FIELDS = ("material_lut_x", "material_lut_y", "ruling_lut_x", "ruling_lut_y")


def table_key(table) -> bytes:
    """
    The content of a lowered table as a hashable key: the packed records with their host pointers
    blanked, followed by the arrays those pointers refer to.  Two lowerings of an unchanged system
    give the same key; any edit of a surface that matters to the trace changes it.
    """
    raw = bytearray(table)
    size = C.sizeof(L.Surface)
    for k in range(len(table)):
        for name in _POINTER_FIELDS:
            field = getattr(L.Surface, name)
            raw[k * size + field.offset: k * size + field.offset + field.size] = bytes(field.size)
    return bytes(raw) + b"".join(np.ascontiguousarray(a).tobytes() for a in getattr(table, "keep", ()))


def _table_type(n: int):
    class SurfaceTable(L.Surface * n):  # a Python subclass, so instances can carry `keep`
        pass

    return SurfaceTable


def fingerprint(obj) -> bytes:
    """
    A digest of everything a surface list is made of: dataclass fields, named arrays (axes and bytes),
    NumPy arrays, scalars, containers -- walked recursively.  Two calls on an unchanged system give the
    same digest at a fraction of the cost of lowering it (which evaluates every parameter at every
    configuration), so ``SequentialSystem`` lowers again only when the digest moves.
    """
    import dataclasses
    import hashlib

    h = hashlib.blake2b(digest_size=16)

    def walk(o, depth=0):
        if depth > 32:
            raise RecursionError("fingerprint: structure too deep")
        if o is None or isinstance(o, (bool, int, float, complex, str, bytes)):
            h.update(repr(o).encode())
        elif isinstance(o, np.ndarray):
            h.update(o.dtype.str.encode())
            h.update(repr(o.shape).encode())
            h.update(np.ascontiguousarray(o).tobytes())
        elif isinstance(o, np.generic):
            h.update(repr(o.item()).encode())
        elif isinstance(o, na.ScalarArray):
            h.update(repr(o.axes).encode())
            walk(np.asarray(o.ndarray), depth + 1)
        elif dataclasses.is_dataclass(o) and not isinstance(o, type):
            h.update(type(o).__qualname__.encode())
            for f in dataclasses.fields(o):
                h.update(f.name.encode())
                walk(getattr(o, f.name), depth + 1)
        elif isinstance(o, dict):
            for k, v in o.items():
                walk(k, depth + 1)
                walk(v, depth + 1)
        elif isinstance(o, (list, tuple)):
            h.update(b"[")
            for v in o:
                walk(v, depth + 1)
            h.update(b"]")
        elif hasattr(o, "__dict__"):
            h.update(type(o).__qualname__.encode())
            walk({k: v for k, v in vars(o).items() if not k.startswith("_")}, depth + 1)
        else:
            h.update(repr(o).encode())

    walk(obj)
    return h.digest()
