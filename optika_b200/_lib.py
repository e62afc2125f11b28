"""
ctypes binding of ``liboptk.so`` (C ABI declared in ``include/optk.h``).

There is deliberately no fallback: if the shared library is missing, or a
compute entry point fails (no CUDA device, bad arguments), the call raises.
Status codes map to the exceptions the reference raises for the same mistakes
(``ValueError`` for bad arguments, e.g. ``optika/apertures/_apertures.py:98-99``).
"""

from __future__ import annotations
import ctypes as C
import os
import pathlib

__all__ = [
    "lib",
    "check",
    "OptkError",
    "Affine",
    "Surface",
    "RaysIn",
    "RaysOut",
    "Image",
    "TraceStats",
    "MlLayer",
    "MlSegment",
    "MlInput",
]

MAX_SURFACES = 22
MAX_VERTICES = 32
MAX_COEFF = 8
MAX_AXES = 8
NUM_FIELDS = 10
ML_MAX_AXES = 4
ML_MAX_LAYERS = 256

# optk_field_t
FIELDS = (
    "wavelength",
    "px", "py", "pz",
    "dx", "dy", "dz",
    "intensity",
    "attenuation",
    "index_refraction",
)

SAG_FLAT, SAG_SPHERICAL, SAG_CYLINDRICAL, SAG_CONIC, SAG_PARABOLIC, SAG_TOROIDAL = range(6)
MAT_VACUUM, MAT_MIRROR, MAT_GLASS, MAT_INDEX, MAT_INDEX_MIRROR, MAT_PASS = range(6)
RULING_NONE, RULING_CONSTANT, RULING_POLYNOMIAL, RULING_HOLOGRAPHIC = range(4)
(
    APERTURE_NONE,
    APERTURE_CIRCULAR,
    APERTURE_RECTANGULAR,
    APERTURE_POLYGON,
    APERTURE_ELLIPTICAL,
    APERTURE_SECTOR,
) = range(6)

F_TRANSFORM = 0x001
F_SAG_TRANSFORM = 0x002
F_APERTURE_TRANSFORM = 0x004
F_RULING_TRANSFORM = 0x008
F_APERTURE_INVERTED = 0x010
F_APERTURE_ACTIVE = 0x020
F_APERTURE_ANGULAR = 0x040
F_HOLO_DIVERGING_1 = 0x080
F_HOLO_DIVERGING_2 = 0x100
F_LOCAL_OUT = 0x200
F_TRANSLATION_ONLY = 0x400  # set by the library
F_APERTURE_CONVEX = 0x4000  # set by the library (optk_system_create)
F_APERTURE_CLOCKWISE = 0x8000

STAGE_INTERCEPT = 0x01
STAGE_ATTENUATE = 0x02
STAGE_RULINGS = 0x04
STAGE_REFRACT = 0x08
STAGE_CLIP = 0x10
STAGE_NORMAL_OUT = 0x20
STAGE_SAG_OUT = 0x40
STAGE_KAPPA_OUT = 0x80
STAGE_EFFICIENCY_OUT = 0x100
STAGE_ALL = 0x1F
EFF_UNIT, EFF_LUT, EFF_TABLE2D = range(3)
(PROFILE_IDEAL, PROFILE_SINUSOIDAL, PROFILE_SQUARE, PROFILE_SAWTOOTH, PROFILE_TRIANGULAR, PROFILE_RECTANGULAR,
 PROFILE_MEASURED) = range(7)


class Affine(C.Structure):
    _fields_ = [("r", C.c_double * 9), ("t", C.c_double * 3)]


class Surface(C.Structure):
    _fields_ = [
        ("sag_kind", C.c_int32),
        ("material_kind", C.c_int32),
        ("ruling_kind", C.c_int32),
        ("aperture_kind", C.c_int32),
        ("flags", C.c_int32),
        ("stages", C.c_int32),
        ("n_vertices", C.c_int32),
        ("n_coeff", C.c_int32),
        ("transform", Affine),
        ("sag_transform", Affine),
        ("aperture_transform", Affine),
        ("ruling_transform", Affine),
        ("sag", C.c_double * 4),
        ("material", C.c_double * 6),
        ("ruling_order", C.c_double),
        ("ruling_normal", C.c_double * 3),
        ("ruling_coeff", C.c_double * MAX_COEFF),
        ("ruling_power", C.c_int32 * MAX_COEFF),
        ("holo_x1", C.c_double * 3),
        ("holo_x2", C.c_double * 3),
        ("holo_wavelength", C.c_double),
        ("aperture", C.c_double * 4),
        ("vertices_x", C.c_double * MAX_VERTICES),
        ("vertices_y", C.c_double * MAX_VERTICES),
        ("material_efficiency", C.c_int32),
        ("ruling_profile", C.c_int32),
        ("material_lut_n", C.c_int32),
        ("ruling_lut_n", C.c_int32),
        ("ruling_depth", C.c_double),
        ("ruling_duty", C.c_double),
        ("material_lut_x", C.c_void_p),
        ("material_lut_y", C.c_void_p),
        ("ruling_lut_x", C.c_void_p),
        ("ruling_lut_y", C.c_void_p),
    ]


class RaysIn(C.Structure):
    _fields_ = [
        ("n_axes", C.c_int32),
        ("dims", C.c_int64 * MAX_AXES),
        ("field", C.c_void_p * NUM_FIELDS),
        ("stride", (C.c_int64 * MAX_AXES) * NUM_FIELDS),
        ("unvignetted", C.c_void_p),
        ("mask_stride", C.c_int64 * MAX_AXES),
        ("normal", C.c_void_p * 3),
        ("normal_stride", (C.c_int64 * MAX_AXES) * 3),
    ]


class RaysOut(C.Structure):
    _fields_ = [("field", C.c_void_p * NUM_FIELDS), ("unvignetted", C.c_void_p), ("cos_incidence", C.c_void_p)]


class Image(C.Structure):
    _fields_ = [
        ("n_wavelength", C.c_int32),
        ("n_x", C.c_int32),
        ("n_y", C.c_int32),
        ("edges_wavelength", C.c_void_p),
        ("edges_x", C.c_void_p),
        ("edges_y", C.c_void_p),
        ("flux", C.c_void_p),
        ("moment_real", C.c_void_p),
        ("moment_imag", C.c_void_p),
        ("counts", C.c_void_p),
        ("has_range", C.c_int32),
        ("uniform_edges", C.c_int32),
        ("range", C.c_double * 6),
        ("group_size", C.c_int64),
    ]


class Grid(C.Structure):
    _fields_ = [
        ("n", C.c_int32 * 5),
        ("begin", C.c_int32 * 5),
        ("count", C.c_int32 * 5),
        ("at_infinity", C.c_int32),
        ("jitter", C.c_int32),
        ("has_frame", C.c_int32),
        ("seed", C.c_uint64),
        ("vertices", C.c_void_p * 5),
        ("weight_scene", C.c_void_p),
        ("weight_pupil", C.c_void_p),
        ("frame", Affine),
        ("field_2d", C.c_int32),
        ("pupil_2d", C.c_int32),
        ("angular_cells", C.c_void_p * 2),
        ("chromatic", C.c_int32),
        ("weight_pupil_chromatic", C.c_int32),
    ]


STOP_DIRECTION, STOP_POSITION = 0, 1


class StopProblem(C.Structure):
    _fields_ = [
        ("variable", C.c_int32),
        ("target", C.c_int32),
        ("surf_first", C.c_int32),
        ("surf_last", C.c_int32),
        ("max_iterations", C.c_int32),
        ("reserved", C.c_int32),
        ("step", C.c_double),
        ("max_abs_error", C.c_double),
    ]


class TraceStats(C.Structure):
    _fields_ = [
        ("n_rays", C.c_uint64),
        ("n_unvignetted", C.c_uint64),
        ("n_newton_iterations", C.c_uint64),
        ("n_binned", C.c_uint64),
    ]


class MlLayer(C.Structure):
    _fields_ = [
        ("n_re", C.c_void_p),
        ("n_im", C.c_void_p),
        ("n_stride", C.c_int64 * ML_MAX_AXES),
        ("thickness", C.c_void_p),
        ("thickness_stride", C.c_int64 * ML_MAX_AXES),
        ("width", C.c_void_p),
        ("width_stride", C.c_int64 * ML_MAX_AXES),
        ("profile_kind", C.c_int32),
        ("reserved", C.c_int32),
    ]


class MlSegment(C.Structure):
    _fields_ = [
        ("first", C.c_int32),
        ("count", C.c_int32),
        ("repeat", C.c_int32),
        ("reserved", C.c_int32),
    ]


class MlInput(C.Structure):
    _fields_ = [
        ("n_axes", C.c_int32),
        ("dims", C.c_int64 * ML_MAX_AXES),
        ("wavelength", C.c_void_p),
        ("wavelength_stride", C.c_int64 * ML_MAX_AXES),
        ("direction_re", C.c_void_p),
        ("direction_im", C.c_void_p),
        ("direction_stride", C.c_int64 * ML_MAX_AXES),
        ("n_re", C.c_void_p),
        ("n_im", C.c_void_p),
        ("n_stride", C.c_int64 * ML_MAX_AXES),
    ]


class CcdPlane(C.Structure):
    _fields_ = [
        ("energy", C.c_double),
        ("absorption", C.c_double),
        ("thickness_implant", C.c_double),
        ("thickness_depletion", C.c_double),
        ("thickness_substrate", C.c_double),
        ("width_pixel_x", C.c_double),
        ("width_pixel_y", C.c_double),
        ("cce_backsurface", C.c_double),
        ("energy_pair_inf", C.c_double),
        ("fano_inf", C.c_double),
        ("n_pmf", C.c_int32),
        ("reserved", C.c_int32),
        ("cmf", C.c_void_p),
        ("n_values", C.c_void_p),
    ]


ABI_VERSION = 9


class OptkError(RuntimeError):
    pass


# OPTK_LIBRARY points experiments at another build of the same ABI
_PATH = pathlib.Path(os.environ.get("OPTK_LIBRARY") or pathlib.Path(__file__).parent / "liboptk.so")

# every symbol include/optk.h declares
SYMBOLS = (
    "optk_abi_version",
    "optk_last_error",
    "optk_device_count",
    "optk_system_create",
    "optk_system_destroy",
    "optk_system_size",
    "optk_system_surface",
    "optk_trace",
    "optk_trace_host",
    "optk_trace_grid",
    "optk_bin",
    "optk_multilayer",
    "optk_solve_stops",
    "optk_reduce_groups",
    "optk_jit_mode",
    "optk_jit_compiled",
    "optk_interp",
    "optk_apply_efficiency",
    "optk_measure_fp64_peak",
    "optk_measure_soa_copy",
    "optk_host_register",
    "optk_host_unregister",
    "optk_memcpy_async",
    "optk_enable_peer_access",
    "optk_electrons_measured",
    "optk_debug_math",
)

_lib = None


def lib() -> C.CDLL:
    """Load ``liboptk.so`` (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _PATH.exists():
        raise OptkError(
            f"{_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C optika_b200/csrc`). There is no CPU fallback."
        )
    L = C.CDLL(str(_PATH))
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.optk_abi_version.restype = C.c_int
    L.optk_last_error.restype = C.c_char_p
    L.optk_device_count.restype = C.c_int
    L.optk_system_create.argtypes = [C.POINTER(Surface), i32, i32, C.POINTER(vp)]
    L.optk_system_destroy.argtypes = [vp]
    L.optk_system_size.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.optk_system_surface.argtypes = [vp, i32, i32, C.POINTER(Surface)]
    L.optk_trace.argtypes = [
        vp, i32, C.POINTER(RaysIn), C.POINTER(RaysOut), i32, i32, i32, i32, i64,
        C.POINTER(Image), C.POINTER(Affine), vp, vp,
    ]
    L.optk_trace_host.argtypes = [
        vp, i32, C.POINTER(RaysIn), C.POINTER(RaysOut), i32, i32, i32, i32, i64,
        C.POINTER(Image), C.POINTER(Affine), C.POINTER(TraceStats), i64, i32,
    ]
    L.optk_trace_grid.argtypes = [
        vp, i32, C.POINTER(Grid), C.POINTER(RaysOut), i32, i32, i32, i32, i64,
        C.POINTER(Image), C.POINTER(Affine), vp, vp,
    ]
    L.optk_bin.argtypes = [i64, vp, vp, vp, vp, vp, vp, C.POINTER(Image), vp]
    L.optk_multilayer.argtypes = [
        C.POINTER(MlInput), i32, C.POINTER(MlLayer), i32, C.POINTER(MlSegment), vp, vp, vp, vp, vp,
    ]
    L.optk_jit_mode.argtypes = [i32]
    L.optk_reduce_groups.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.optk_solve_stops.argtypes = [vp, i32, C.POINTER(StopProblem), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.optk_interp.argtypes = [i64, vp, i32, vp, vp, vp, vp, vp, vp]
    L.optk_apply_efficiency.argtypes = [i64, vp, vp, vp, vp]
    L.optk_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), vp]
    L.optk_measure_soa_copy.argtypes = [i64, C.POINTER(C.c_double), vp]
    L.optk_host_register.argtypes = [vp, i64]
    L.optk_host_unregister.argtypes = [vp]
    L.optk_memcpy_async.argtypes = [vp, vp, i64, vp]
    L.optk_enable_peer_access.argtypes = [i32]
    L.optk_debug_math.argtypes = [i32, i64, vp, vp, vp, vp]
    L.optk_electrons_measured.argtypes = [i32, i32, i32, C.POINTER(CcdPlane), vp, vp, i32, C.c_uint64, vp]
    for name in SYMBOLS:
        if name not in ("optk_last_error",):
            getattr(L, name).restype = C.c_int
    L.optk_jit_compiled.restype = C.c_int64
    if L.optk_abi_version() != ABI_VERSION:
        raise OptkError("liboptk.so ABI version mismatch; rebuild it")
    _lib = L
    return L


def check(status: int) -> int:
    """Raise the Python exception matching a negative ``optk_status_t``."""
    if status >= 0:
        return status
    message = lib().optk_last_error().decode("utf-8", "replace")
    if status == -1:
        raise ValueError(message)
    if status == -2:
        raise NotImplementedError(message)
    if status == -4:
        raise MemoryError(message)
    raise OptkError(message)
