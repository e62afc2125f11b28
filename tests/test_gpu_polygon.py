"""
Polygon apertures: convex polygons are decided by half-plane tests in contracted arithmetic, with the reference's
operation-by-operation even-odd arithmetic only inside a 1e-12 B^2 band around the edge lines
(``OPTK_F_APERTURE_CONVEX``, ``csrc/trace_impl.cuh::aperture_test``).  The decisions must equal the oracle's
(``oracle/raytrace.py``, na.geometry.point_in_polygon restated) for EVERY point, including the ones exactly on
vertices and edges, one rounding away from them, on the extension of an edge line, and far outside; with both the
table-driven and the run-time specialised kernels; for clockwise and counter-clockwise vertex orders; and polygons that
are not convex must keep the even-odd rule.
"""

import numpy as np
import pytest

import optika_b200 as optika
from optika_b200 import named as na
from optika_b200 import units as u
from optika_b200 import _lib, _lowering
from oracle import raytrace as ora

import configs
from test_gpu_trace import host_states

pytestmark = pytest.mark.gpu


def polygon(x, y, **kwargs):
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    return optika.apertures.PolygonalAperture(
        vertices=na.Cartesian3dVectorArray(na.ScalarArray(x, "vertex"), na.ScalarArray(y, "vertex"), na.ScalarArray(0 * x, "vertex")),
        **kwargs,
    )


def regular(n, radius, phase=0.1, clockwise=False):
    a = phase + 2 * np.pi * np.arange(n) / n
    if clockwise:
        a = a[::-1]
    return radius * np.cos(a), radius * np.sin(a)


PENTAGRAM = tuple(np.array(regular(5, 12.0))[:, [0, 2, 4, 1, 3]])  # convex position, but the edges cross: winds twice

POLYGONS = {
    "octagon": (lambda: optika.apertures.OctagonalAperture(14.0), True),
    "pentagon_ccw": (lambda: polygon(*regular(5, 13.0)), True),
    "heptagon_clockwise": (lambda: polygon(*regular(7, 13.0, clockwise=True)), True),
    "hexagon_inverted": (lambda: optika.apertures.RegularPolygonalAperture(14.0, 6, inverted=True), True),
    "trapezoid": (lambda: optika.apertures.IsoscelesTrapezoidalAperture(x_left=3.0, x_right=16.0, angle=50 * u.deg), True),
    "sixteen": (lambda: polygon(*regular(16, 15.0)), True),
    "thirty_two": (lambda: polygon(*regular(32, 15.0, phase=0.03)), True),  # OPTK_MAX_VERTICES
    "star_of_twenty": (lambda: polygon(*(np.array(regular(20, 1.0)) * np.where(np.arange(20) % 2, 6.0, 14.0))), False),
    "triangle_far_from_origin": (lambda: polygon([100.0, 130.0, 110.0], [200.0, 205.0, 240.0]), True),
    "arrow_not_convex": (lambda: polygon([-10.0, 12.0, 4.0, 9.0, -6.0], [-8.0, -9.0, 0.0, 11.0, 7.0]), False),
    "pentagram": (lambda: polygon(*PENTAGRAM), False),
    "collinear_vertex": (lambda: polygon([-10.0, 0.0, 10.0, 10.0, -10.0], [-5.0, -5.0, -5.0, 5.0, 5.0]), False),
}


def probe_points(vx, vy, seed=0):
    """Random points plus the ones where the decision hangs on one rounding."""
    rng = np.random.default_rng(seed)
    cx, cy = vx.mean(), vy.mean()
    span = max(np.ptp(vx), np.ptp(vy))
    x = [cx + span * rng.uniform(-0.8, 0.8, 20000)]
    y = [cy + span * rng.uniform(-0.8, 0.8, 20000)]
    n = len(vx)
    for i in range(n):
        j = (i + 1) % n
        for t in (0.0, 0.25, 0.5, 1.0 / 3.0, 0.875, 1.0, -0.5, 1.5, 40.0):  # on the segment, and on its extension
            px, py = vx[i] + t * (vx[j] - vx[i]), vy[i] + t * (vy[j] - vy[i])
            for k in (-2, -1, 0, 1, 2):  # ... and a few roundings to either side
                x.append(np.array([px + k * np.spacing(abs(px) + 1e-300), px, px + k * np.spacing(abs(px) + 1e-300)]))
                y.append(np.array([py, py + k * np.spacing(abs(py) + 1e-300), py - k * np.spacing(abs(py) + 1e-300)]))
        # points a 1e-13 th of the polygon's size from the edge: inside the band, not on the line
        mx, my = 0.5 * (vx[i] + vx[j]), 0.5 * (vy[i] + vy[j])
        nx, ny = vy[j] - vy[i], -(vx[j] - vx[i])
        for eps in (1e-13, -1e-13, 1e-10, -1e-10, 1e-7, -1e-7):
            x.append(np.array([mx + eps * nx]))
            y.append(np.array([my + eps * ny]))
    x.append(np.array([1e6, -1e6, 0.0, np.nan, 1.0, np.inf, cx]))
    y.append(np.array([0.0, 1e6, -1e9, 1.0, np.nan, 0.0, cy]))
    return np.concatenate(x), np.concatenate(y)


def vertices_of(aperture):
    v = aperture.vertices
    return np.asarray(na.as_named_array(v.x).ndarray, dtype=float).ravel(), np.asarray(na.as_named_array(v.y).ndarray, dtype=float).ravel()


def rays_at(x, y, repeat=1):
    x, y = np.tile(x, repeat), np.tile(y, repeat)
    return optika.rays.RayVectorArray(
        wavelength=5e-4,
        position=na.Cartesian3dVectorArray(na.ScalarArray(x, "ray"), na.ScalarArray(y, "ray"), -1.0),
        direction=na.Cartesian3dVectorArray(0.0, 0.0, 1.0),
    )


@pytest.mark.parametrize("name", list(POLYGONS))
def test_polygon_masks_equal_the_oracle_everywhere(cuda_device, name):
    make, convex = POLYGONS[name]
    aperture = make()
    vx, vy = vertices_of(aperture)
    x, y = probe_points(vx, vy)
    rays = rays_at(x, y)
    surface = optika.surfaces.Surface(aperture=aperture)
    table, _ = _lowering.lower_system([surface], stages=_lib.STAGE_ALL)
    r0, _ = configs.flatten_rays(rays)
    want = ora.surface_propagate(surface, r0)["unvignetted"]
    got = host_states(surface.propagate_rays(rays))["unvignetted"]
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]
    assert 0.05 < want[:20000].mean() < 0.95  # the random part samples both sides


@pytest.mark.parametrize("name", ["octagon", "heptagon_clockwise", "arrow_not_convex", "triangle_far_from_origin", "thirty_two", "star_of_twenty"])
def test_polygon_masks_in_the_specialised_kernel(cuda_device, name):
    """The same points through the kernel NVRTC compiles for the surface (long launch, OPTK_JIT forced on)."""
    lib = _lib.lib()
    make, convex = POLYGONS[name]
    aperture = make()
    vx, vy = vertices_of(aperture)
    x, y = probe_points(vx, vy, seed=3)
    # (no surface transformation: the oracle applies every transformation as a full matrix, which turns a NaN x
    # into a NaN y as well; a translation in the reference and in the kernels does not)
    surface = optika.surfaces.Surface(aperture=aperture)
    rays = rays_at(x, y, repeat=8)
    r0, _ = configs.flatten_rays(rays)
    want = ora.surface_propagate(surface, r0)["unvignetted"]
    before = lib.optk_jit_compiled()
    try:
        _lib.check(lib.optk_jit_mode(1))
        got = host_states(surface.propagate_rays(rays))["unvignetted"]
    finally:
        _lib.check(lib.optk_jit_mode(-1))
    assert lib.optk_jit_compiled() >= before
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]


def test_convexity_is_classified_by_the_library(cuda_device):
    """
    optk_system_create marks strictly convex vertex lists (and their orientation) and nothing else: read back
    through the masks -- a pentagram's centre pentagon is OUTSIDE under the even-odd rule (two crossings), which
    the half-plane test of a wrongly classified polygon would get wrong.
    """
    aperture = polygon(*PENTAGRAM)
    surface = optika.surfaces.Surface(aperture=aperture)
    rays = rays_at(np.array([0.0, 11.0 * np.cos(0.1), 30.0]), np.array([0.0, 11.0 * np.sin(0.1), 0.0]))
    got = host_states(surface.propagate_rays(rays))["unvignetted"]
    assert got.tolist() == [False, True, False]


def test_more_vertices_than_the_table_holds_are_refused():
    with pytest.raises(ValueError, match="at most"):
        surface = optika.surfaces.Surface(aperture=polygon(*regular(_lib.MAX_VERTICES + 1, 10.0)))
        _lowering.lower_system([surface], stages=_lib.STAGE_ALL)
