"""
The C-ABI library loads on a CPU-only box, exports every symbol that
``include/optk.h`` declares, agrees with the ctypes mirror on struct layout, and
fails loudly (no CPU fallback) when there is no CUDA device.
"""

import ctypes as C
import pathlib
import re
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "optk.h"


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    if not (ROOT / "optika_b200" / "liboptk.so").exists():
        __graft_entry__.build()
    from optika_b200 import _lib

    return _lib.lib()


def declared_symbols():
    text = HEADER.read_text()
    return re.findall(r"OPTK_API\s+[\w\s\*]+?\b(optk_\w+)\s*\(", text)


def test_exports_every_declared_symbol(lib):
    from optika_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 11
    assert set(names) == set(_lib.SYMBOLS)
    for name in names:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", str(ROOT / "optika_b200" / "liboptk.so")], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (optk_\w+)", out))
    assert set(names) <= exported


def test_struct_layout_matches_c(tmp_path, lib):
    from optika_b200 import _lib

    structs = {
        "optk_affine_t": _lib.Affine,
        "optk_surface_t": _lib.Surface,
        "optk_rays_in_t": _lib.RaysIn,
        "optk_rays_out_t": _lib.RaysOut,
        "optk_image_t": _lib.Image,
        "optk_trace_stats_t": _lib.TraceStats,
        "optk_grid_t": _lib.Grid,
        "optk_ml_layer_t": _lib.MlLayer,
        "optk_ml_segment_t": _lib.MlSegment,
        "optk_ml_input_t": _lib.MlInput,
        "optk_stop_problem_t": _lib.StopProblem,
        "optk_ccd_plane_t": _lib.CcdPlane,
    }
    probes = [
        ("optk_surface_t", "transform"), ("optk_surface_t", "sag"), ("optk_surface_t", "ruling_power"),
        ("optk_surface_t", "holo_wavelength"), ("optk_surface_t", "vertices_y"),
        ("optk_rays_in_t", "field"), ("optk_rays_in_t", "stride"), ("optk_rays_in_t", "mask_stride"),
        ("optk_rays_in_t", "normal_stride"), ("optk_image_t", "edges_wavelength"), ("optk_image_t", "counts"), ("optk_image_t", "range"),
        ("optk_grid_t", "seed"), ("optk_grid_t", "vertices"), ("optk_grid_t", "frame"),
        ("optk_ml_layer_t", "width_stride"), ("optk_ml_layer_t", "profile_kind"),
        ("optk_ml_input_t", "direction_stride"), ("optk_ml_input_t", "n_stride"),
        ("optk_image_t", "group_size"), ("optk_stop_problem_t", "max_iterations"), ("optk_stop_problem_t", "step"),
        ("optk_stop_problem_t", "max_abs_error"),
        ("optk_image_t", "uniform_edges"), ("optk_grid_t", "angular_cells"), ("optk_grid_t", "chromatic"),
        ("optk_grid_t", "weight_pupil_chromatic"), ("optk_ccd_plane_t", "fano_inf"), ("optk_ccd_plane_t", "n_pmf"),
        ("optk_ccd_plane_t", "cmf"), ("optk_ccd_plane_t", "n_values"),
    ]
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "optk.h"', "int main(void) {"]
    for name in structs:
        src.append(f'printf("sizeof {name} %zu\\n", sizeof({name}));')
    for s, f in probes:
        src.append(f'printf("offsetof {s} {f} %zu\\n", offsetof({s}, {f}));')
    src += ["return 0;", "}"]
    c_file = tmp_path / "layout.c"
    c_file.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(c_file), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        parts = line.split()
        if parts[0] == "sizeof":
            assert C.sizeof(structs[parts[1]]) == int(parts[2]), line
        else:
            assert getattr(structs[parts[1]], parts[2]).offset == int(parts[3]), line


def test_constants_match_header():
    from optika_b200 import _lib

    text = HEADER.read_text()

    def define(name):
        return int(re.search(rf"#define {name} (\w+)", text).group(1), 0)

    assert define("OPTK_MAX_SURFACES") == _lib.MAX_SURFACES
    assert define("OPTK_MAX_VERTICES") == _lib.MAX_VERTICES
    assert define("OPTK_MAX_COEFF") == _lib.MAX_COEFF
    assert define("OPTK_MAX_AXES") == _lib.MAX_AXES
    assert define("OPTK_ML_MAX_AXES") == _lib.ML_MAX_AXES
    assert define("OPTK_STAGE_ALL") == _lib.STAGE_ALL
    assert define("OPTK_F_LOCAL_OUT") == _lib.F_LOCAL_OUT
    assert define("OPTK_F_TRANSLATION_ONLY") == _lib.F_TRANSLATION_ONLY
    assert define("OPTK_F_APERTURE_CONVEX") == _lib.F_APERTURE_CONVEX
    assert define("OPTK_F_APERTURE_CLOCKWISE") == _lib.F_APERTURE_CLOCKWISE
    assert define("OPTK_STAGE_KAPPA_OUT") == _lib.STAGE_KAPPA_OUT


def test_system_create_validates_without_a_gpu(lib):
    from optika_b200 import _lib

    table = (_lib.Surface * 1)()
    table[0].sag_kind = 99
    handle = C.c_void_p()
    with pytest.raises(NotImplementedError):
        _lib.check(lib.optk_system_create(table, 1, 1, C.byref(handle)))
    table[0].sag_kind = _lib.SAG_FLAT
    table[0].aperture_kind = _lib.APERTURE_POLYGON
    table[0].n_vertices = 2
    with pytest.raises(ValueError):
        _lib.check(lib.optk_system_create(table, 1, 1, C.byref(handle)))
    table[0].aperture_kind = _lib.APERTURE_NONE
    _lib.check(lib.optk_system_create(table, 1, 1, C.byref(handle)))
    n_s, n_c = C.c_int32(), C.c_int32()
    _lib.check(lib.optk_system_size(handle, C.byref(n_s), C.byref(n_c)))
    assert (n_s.value, n_c.value) == (1, 1)
    _lib.check(lib.optk_system_destroy(handle))


def test_what_the_library_derives_per_surface(lib):
    """
    optk_system_create adds per-surface constants and flags (read back with optk_system_surface, no GPU): which
    polygons are strictly convex and how they are oriented -- the half-plane test of the kernels relies on it --,
    the squared-radius threshold of circles, rotation-free frames.
    """
    import numpy as np

    from optika_b200 import _lib

    def regular(n, radius=10.0, phase=0.1):
        a = phase + 2 * np.pi * np.arange(n) / n
        return radius * np.cos(a), radius * np.sin(a)

    pentagram = tuple(np.array(regular(5))[:, [0, 2, 4, 1, 3]])
    cases = {
        "octagon": (regular(8), _lib.F_APERTURE_CONVEX),
        "clockwise": (tuple(v[::-1] for v in regular(7)), _lib.F_APERTURE_CONVEX | _lib.F_APERTURE_CLOCKWISE),
        "thirty_two": (regular(_lib.MAX_VERTICES), _lib.F_APERTURE_CONVEX),
        "far_from_origin": (([100.0, 130.0, 110.0], [200.0, 205.0, 240.0]), _lib.F_APERTURE_CONVEX),
        "arrow": (([-10.0, 12.0, 4.0, 9.0, -6.0], [-8.0, -9.0, 0.0, 11.0, 7.0]), 0),
        "pentagram": (pentagram, 0),  # every turn has the same sign, but it winds twice
        "collinear": (([-10.0, 0.0, 10.0, 10.0, -10.0], [-5.0, -5.0, -5.0, 5.0, 5.0]), 0),
        "repeated": (([-10.0, 10.0, 10.0, 10.0, -10.0], [-5.0, -5.0, -5.0, 5.0, 5.0]), 0),
        "degenerate": (([0.0, 1.0, 2.0], [0.0, 1.0, 2.0]), 0),
    }
    table = (_lib.Surface * (len(cases) + 2))()
    for k, ((vx, vy), _) in enumerate(cases.values()):
        S = table[k]
        S.sag_kind, S.aperture_kind, S.stages = _lib.SAG_FLAT, _lib.APERTURE_POLYGON, _lib.STAGE_ALL
        S.flags = _lib.F_APERTURE_ACTIVE | _lib.F_APERTURE_CONVEX  # a caller cannot claim convexity
        S.n_vertices = len(vx)
        for i in range(len(vx)):
            S.vertices_x[i], S.vertices_y[i] = float(vx[i]), float(vy[i])
    circle, shifted = table[len(cases)], table[len(cases) + 1]
    circle.sag_kind, circle.aperture_kind, circle.stages = _lib.SAG_SPHERICAL, _lib.APERTURE_CIRCULAR, _lib.STAGE_ALL
    circle.sag[0], circle.aperture[0] = 250.0, 12.3
    shifted.sag_kind, shifted.stages, shifted.flags = _lib.SAG_FLAT, _lib.STAGE_ALL, _lib.F_TRANSFORM
    for i in (0, 4, 8):
        shifted.transform.r[i] = 1.0
    shifted.transform.t[2] = 40.0
    for S in table:
        for i in (0, 4, 8):
            S.sag_transform.r[i] = S.aperture_transform.r[i] = S.ruling_transform.r[i] = 1.0
            if not S.flags & _lib.F_TRANSFORM:
                S.transform.r[i] = 1.0
    handle = C.c_void_p()
    _lib.check(lib.optk_system_create(table, len(table), 1, C.byref(handle)))
    try:
        got = _lib.Surface()
        for k, (name, ((vx, vy), want)) in enumerate(cases.items()):
            _lib.check(lib.optk_system_surface(handle, 0, k, C.byref(got)))
            mask = _lib.F_APERTURE_CONVEX | _lib.F_APERTURE_CLOCKWISE
            assert got.flags & mask == want, name
            if want:
                bound = max(np.abs(vx).max(), np.abs(vy).max())
                assert got.aperture[0] == bound and np.isclose(got.aperture[1], 1e-12 * bound**2, rtol=1e-12), name
            assert list(got.vertices_x[: len(vx)]) == [float(v) for v in vx]  # the vertices themselves are untouched
        _lib.check(lib.optk_system_surface(handle, 0, len(cases), C.byref(got)))
        threshold = got.aperture[3]
        assert np.sqrt(threshold) <= 12.3 < np.sqrt(np.nextafter(threshold, np.inf)) and got.sag[3] == 1 / 250.0
        _lib.check(lib.optk_system_surface(handle, 0, len(cases) + 1, C.byref(got)))
        assert got.flags & _lib.F_TRANSLATION_ONLY and got.transform.t[2] == 40.0
        with pytest.raises(ValueError):
            _lib.check(lib.optk_system_surface(handle, 1, 0, C.byref(got)))
        with pytest.raises(ValueError):
            _lib.check(lib.optk_system_surface(handle, 0, len(table), C.byref(got)))
    finally:
        _lib.check(lib.optk_system_destroy(handle))


def test_polygon_classification_on_random_polygons(lib):
    """
    The library's convexity test against an independent criterion on 300 random vertex lists: points on a circle in
    angular order (convex, either orientation), the same with one vertex pulled inside (not convex), and random
    permutations (edges cross).  Independent criterion: every vertex strictly inside every non-adjacent edge's
    half-plane, by exact rational arithmetic.
    """
    import fractions
    import numpy as np

    from optika_b200 import _lib

    def convex_exact(vx, vy):
        fx, fy = [fractions.Fraction(float(v)) for v in vx], [fractions.Fraction(float(v)) for v in vy]
        n = len(fx)
        area2 = sum(fx[i] * fy[(i + 1) % n] - fx[(i + 1) % n] * fy[i] for i in range(n))
        if area2 == 0:
            return 0, 0
        o = 1 if area2 > 0 else -1
        margin = min(
            o * ((fx[(i + 1) % n] - fx[i]) * (fy[k] - fy[i]) - (fy[(i + 1) % n] - fy[i]) * (fx[k] - fx[i]))
            for i in range(n) for k in range(n) if k not in (i, (i + 1) % n)
        )
        return (o if margin > 0 else 0), float(margin)

    rng = np.random.default_rng(11)
    polygons = []
    for trial in range(300):
        n = int(rng.integers(3, _lib.MAX_VERTICES + 1))
        angles = np.sort(rng.uniform(0, 2 * np.pi, n))
        if np.min(np.diff(np.concatenate([angles, [angles[0] + 2 * np.pi]]))) < 1e-3:
            continue
        radius = rng.uniform(0.1, 50.0)
        vx, vy = radius * np.cos(angles) + rng.normal(0, 5), radius * np.sin(angles) + rng.normal(0, 5)
        kind = trial % 4
        if kind == 1:
            vx, vy = vx[::-1].copy(), vy[::-1].copy()
        elif kind == 2 and n > 3:
            k = int(rng.integers(n))
            vx[k], vy[k] = 0.5 * (vx[k - 1] + vx[(k + 1) % n]) * 0.9 + 0.1 * vx.mean(), 0.5 * (vy[k - 1] + vy[(k + 1) % n]) * 0.9 + 0.1 * vy.mean()
        elif kind == 3 and n > 4:
            order = rng.permutation(n)
            vx, vy = vx[order], vy[order]
        polygons.append((vx, vy))
    table = (_lib.Surface * len(polygons))()
    for S, (vx, vy) in zip(table, polygons):
        S.sag_kind, S.aperture_kind, S.stages, S.flags = _lib.SAG_FLAT, _lib.APERTURE_POLYGON, _lib.STAGE_ALL, _lib.F_APERTURE_ACTIVE
        S.n_vertices = len(vx)
        for i in range(len(vx)):
            S.vertices_x[i], S.vertices_y[i] = float(vx[i]), float(vy[i])
        for i in (0, 4, 8):
            S.transform.r[i] = S.sag_transform.r[i] = S.aperture_transform.r[i] = S.ruling_transform.r[i] = 1.0
    handle = C.c_void_p()
    _lib.check(lib.optk_system_create(table, len(table), 1, C.byref(handle)))
    try:
        got = _lib.Surface()
        seen = {0: 0, 1: 0, -1: 0}
        for k, (vx, vy) in enumerate(polygons):
            _lib.check(lib.optk_system_surface(handle, 0, k, C.byref(got)))
            orientation, margin = convex_exact(vx, vy)
            bound = max(np.abs(vx).max(), np.abs(vy).max())
            convex = bool(got.flags & _lib.F_APERTURE_CONVEX)
            if abs(margin) > 2e-9 * bound**2:  # the library wants a clear margin (1e-9 B^2); right at it either answer is fine
                assert convex == (orientation != 0), (k, margin, bound)
            if convex:
                assert orientation != 0 and bool(got.flags & _lib.F_APERTURE_CLOCKWISE) == (orientation < 0), k
                seen[orientation] += 1
            else:
                seen[0] += 1
        assert min(seen.values()) > 20  # all three outcomes occur
    finally:
        _lib.check(lib.optk_system_destroy(handle))


def test_stop_solver_and_reductions_validate_their_arguments_without_a_gpu(lib):
    """Argument checks of optk_solve_stops / optk_reduce_groups come before any CUDA call."""
    from optika_b200 import _lib

    table = (_lib.Surface * 3)()
    for k in range(3):
        table[k].sag_kind = _lib.SAG_FLAT
        table[k].stages = _lib.STAGE_ALL
    handle = C.c_void_p()
    _lib.check(lib.optk_system_create(table, 3, 1, C.byref(handle)))
    one = (C.c_double * 1)(0.0)
    counter = (C.c_uint32 * 1)(0)
    ptr = C.cast(one, C.c_void_p)

    def solve(problem, n=1):
        return lib.optk_solve_stops(
            handle, 0, C.byref(problem) if problem is not None else None, n, ptr, ptr, ptr, ptr, ptr, ptr, ptr, ptr, ptr,
            C.cast(counter, C.c_void_p), None,
        )

    good = dict(variable=_lib.STOP_DIRECTION, target=_lib.STOP_POSITION, surf_first=0, surf_last=2, max_iterations=100,
                reserved=0, step=1e-6, max_abs_error=1e-9)
    with pytest.raises(ValueError):
        _lib.check(solve(None))
    for bad in (dict(variable=7), dict(surf_last=0), dict(surf_last=40), dict(step=0.0), dict(max_iterations=0),
                dict(max_abs_error=-1.0)):
        with pytest.raises(ValueError):
            _lib.check(solve(_lib.StopProblem(**{**good, **bad})))
    with pytest.raises(ValueError):
        _lib.check(solve(_lib.StopProblem(**{**good, "surf_first": 1, "surf_last": 3})))  # past the last surface
    assert solve(_lib.StopProblem(**good), n=0) == 0  # nothing to do: no launch
    _lib.check(lib.optk_system_destroy(handle))

    def reduce(n_groups, n_inner, x=ptr):
        return lib.optk_reduce_groups(n_groups, n_inner, x, ptr, None, None, None, None, None, None, None, None, None)

    for n_groups, n_inner in ((1, 0), (-1, 4), (2**20, 2**20)):
        with pytest.raises(ValueError):
            _lib.check(reduce(n_groups, n_inner))
    with pytest.raises(ValueError):
        _lib.check(reduce(1, 1, x=None))
    assert reduce(0, 5) == 0


def test_no_cpu_fallback():
    """Without CUDA the product path raises; it never routes to the oracle or any CPU code."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import optika_b200 as optika
    from optika_b200 import _lib
    import configs

    system = configs.newtonian(num_field=1, num_pupil=2)
    with pytest.raises(_lib.OptkError):
        system.raytrace(**configs.PHYSICAL)
    with pytest.raises(_lib.OptkError):
        optika.materials.multilayer_efficiency(1e-5, 1, 1, [optika.materials.Layer("Si", thickness=1e-5)])
    # the product package must not import the oracle
    import subprocess, sys

    code = "import sys, optika_b200, optika_b200.systems, optika_b200.sensors; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT)).stdout.strip()
    assert out == "False"


def test_round_two_entry_points_validate_their_arguments_without_a_gpu(lib):
    """Efficiency tables, the electron kernel and the host-memory helpers check their arguments before any CUDA call."""
    from optika_b200 import _lib

    table = (_lib.Surface * 1)()
    table[0].sag_kind = _lib.SAG_FLAT
    table[0].stages = _lib.STAGE_ALL
    table[0].material_kind = _lib.MAT_MIRROR
    table[0].material_efficiency = _lib.EFF_TABLE2D  # no table pointers, no node counts
    handle = C.c_void_p()
    with pytest.raises(ValueError, match="2-D efficiency table"):
        _lib.check(lib.optk_system_create(table, 1, 1, C.byref(handle)))
    table[0].material_kind = _lib.MAT_GLASS  # a table needs a mirror or a pass-through material
    with pytest.raises(ValueError, match="mirror or pass-through"):
        _lib.check(lib.optk_system_create(table, 1, 1, C.byref(handle)))
    table[0].material_efficiency = 7
    with pytest.raises(NotImplementedError):
        _lib.check(lib.optk_system_create(table, 1, 1, C.byref(handle)))
    with pytest.raises(ValueError):
        _lib.check(lib.optk_electrons_measured(1, 2, 2, None, None, None, 0, 0, None))
    with pytest.raises(ValueError):
        _lib.check(lib.optk_host_register(None, 16))
    with pytest.raises(ValueError):
        _lib.check(lib.optk_host_unregister(None))
    with pytest.raises(ValueError):
        _lib.check(lib.optk_memcpy_async(None, None, 8, None))
