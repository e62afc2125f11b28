"""
ORACLE (test infrastructure, not the product): CPU restatement of the reference's input
grid for ``SequentialSystem.image`` — cell vertices -> stratified random cell samples,
5-D cell areas, input rays.

Follows
  * ``optika/systems/_sequential.py:1055-1086`` (``_rayfunction_from_vertices``:
    ``grid.broadcast_to(shape).cell_centers(axis=..., random=True)``, ``flux = radiance * area``),
  * ``optika/vectors/_vectors_object.py:42-133`` (``cell_area``),
  * ``optika/systems/_sequential.py:791-828`` (``_calc_rayfunction_input``) and
    ``optika/_util.py:41-73`` (``optika.direction``).

Third-party arithmetic (``named-arrays ~= 2.1``, absent from /root/reference) restated from
its published behaviour; **parity unpinned** for all three (no reference test pins values):
  * ``cell_centers(axis, random=True)``: a uniform random point of every cell, an
    independent draw per axis and per element of the broadcast grid.  The reference draws
    from NumPy's global generator, so no implementation can reproduce its stream; the
    contract here is the *distribution* (one sample per cell, uniform inside it) and a
    documented counter-based stream: Philox4x32-10 (Salmon et al., SC'11; Random123
    known-answer vectors in ``tests/test_oracle_grid.py``), counter = cell index,
    key = seed, five 25-bit integers ``b`` cut from the 128 output bits,
    ``t = (b + 1/2) 2^-25`` (see ``include/optk.h``).
  * ``volume_cell(axis)``: scalar -> difference along the axis; 2-D vector over two axes ->
    signed area of the vertex quadrilateral (half the cross product of its diagonals).
  * ``solid_angle_cell(axis)``: solid angle of the spherical quadrilateral spanned by the
    four direction vertices = two spherical triangles (Van Oosterom & Strackee 1983).
"""

from __future__ import annotations
import numpy as np

__all__ = [
    "philox4x32_10",
    "jitter",
    "cell_samples",
    "direction",
    "input_rays",
    "volume_cell_1d",
    "volume_cell_2d",
    "solid_angle_cell",
    "cell_area",
]

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint64(0x9E3779B9)
_W1 = np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(counter, key):
    """
    Philox4x32 with 10 rounds.  ``counter``: 4 arrays of 32-bit words, ``key``: 2 words.
    Returns 4 ``uint64`` arrays holding 32-bit words.
    """
    c = [np.asarray(v, dtype=np.uint64) & _MASK for v in counter]
    c = list(np.broadcast_arrays(*c))
    k0 = np.uint64(key[0]) & _MASK
    k1 = np.uint64(key[1]) & _MASK
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c


def jitter(cell: np.ndarray, seed: int) -> np.ndarray:
    """``t[5, ...]`` in (0, 1) for the cells with whole-grid C-order index ``cell`` (include/optk.h)."""
    cell = np.asarray(cell, dtype=np.uint64)
    lo, hi = cell & _MASK, cell >> _S32
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    zero = np.zeros_like(lo)
    x = philox4x32_10((lo, hi, zero, zero), key)
    m7, m4 = np.uint64(127), np.uint64(15)
    low = (x[0] & m7) | ((x[1] & m7) << np.uint64(7)) | ((x[2] & m7) << np.uint64(14)) | ((x[3] & m4) << np.uint64(21))
    words = [x[0] >> np.uint64(7), x[1] >> np.uint64(7), x[2] >> np.uint64(7), x[3] >> np.uint64(7), low]
    return np.stack([(w.astype(np.float64) + 0.5) * 2.0**-25 for w in words])


def _cells(vertices, chromatic=()):
    """
    Cells per axis of a grid whose field / pupil vertices are 1-D (separable), 2-D (curvilinear), or --
    for the axes listed in `chromatic` -- 2-D ``[n_wavelength + 1][n_a + 1]``: one row of vertices per
    WAVELENGTH vertex (stop solutions that depend on the wavelength, ``_sequential.py:748-789``).
    """
    v = [np.asarray(a, dtype=np.float64) for a in vertices]

    def cells(a, b):
        if a in chromatic or b in chromatic:
            return [(v[c].shape[1] if c in chromatic else len(v[c])) - 1 for c in (a, b)]
        if v[a].ndim == 2:
            return [v[a].shape[0] - 1, v[a].shape[1] - 1]
        return [len(v[a]) - 1, len(v[b]) - 1]

    return v, [len(v[0]) - 1, *cells(1, 2), *cells(3, 4)]


def cell_samples(vertices, begin=None, count=None, random: bool = True, seed: int = 0, chromatic=()):
    """
    One sample per cell of the sub-box ``[begin, begin + count)`` of a 5-axis vertex grid
    (wavelength, field_x, field_y, pupil_x, pupil_y).  Field and pupil vertices are either
    separable (two 1-D arrays) or curvilinear (two 2-D arrays ``[n_a + 1][n_b + 1]``, sampled
    bilinearly: ``cell_centers`` applied along one axis after the other).  Returns 5 arrays
    of the sub-box shape and the whole-grid cell indices.
    """
    v, n = _cells(vertices, chromatic)
    begin = [0] * 5 if begin is None else list(begin)
    count = [n[a] - begin[a] for a in range(5)] if count is None else list(count)
    idx = np.meshgrid(*[np.arange(begin[a], begin[a] + count[a], dtype=np.int64) for a in range(5)], indexing="ij")
    cell = idx[0].astype(np.uint64)
    for a in range(1, 5):
        cell = cell * np.uint64(n[a]) + idx[a].astype(np.uint64)
    t = jitter(cell, seed) if random else np.full((5,) + cell.shape, 0.5)

    def lerp(lo, hi, ta):
        # the device uses fused multiply-adds here; the difference (<= 1 ulp of the sample) is
        # far below the parity tolerance
        return lo + ta * (hi - lo) if random else 0.5 * (lo + hi)

    out = [lerp(v[0][idx[0]], v[0][idx[0] + 1], t[0])]
    def sample_1d(a):
        if a in chromatic:  # bilinear in (wavelength, axis a): along the wavelength first, as the kernel does
            lo = lerp(v[a][idx[0], idx[a]], v[a][idx[0] + 1, idx[a]], t[0])
            hi = lerp(v[a][idx[0], idx[a] + 1], v[a][idx[0] + 1, idx[a] + 1], t[0])
            return lerp(lo, hi, t[a])
        return lerp(v[a][idx[a]], v[a][idx[a] + 1], t[a])

    for a, b in ((1, 2), (3, 4)):
        if a in chromatic or b in chromatic:
            out.append(sample_1d(a))
            out.append(sample_1d(b))
        elif v[a].ndim == 2:
            for comp in (v[a], v[b]):
                lo = lerp(comp[idx[a], idx[b]], comp[idx[a] + 1, idx[b]], t[a])
                hi = lerp(comp[idx[a], idx[b] + 1], comp[idx[a] + 1, idx[b] + 1], t[a])
                out.append(lerp(lo, hi, t[b]))
        else:
            out.append(lerp(v[a][idx[a]], v[a][idx[a] + 1], t[a]))
            out.append(lerp(v[b][idx[b]], v[b][idx[b] + 1], t[b]))
    return out, idx


def direction(ax, ay):
    """``optika.direction`` (``optika/_util.py:64-73``): angles [rad] -> direction cosines."""
    return -np.cos(ay) * np.sin(ax), -np.sin(ay), np.cos(ay) * np.cos(ax)


def input_rays(
    vertices,
    at_infinity: bool = True,
    weight_scene=None,
    weight_pupil=None,
    begin=None,
    count=None,
    random: bool = True,
    seed: int = 0,
    frame=None,
    chromatic=(),
):
    """
    Flat ray state (dict of 1-D arrays, C order over the sub-box) in the layout of
    ``oracle.raytrace`` (``_sequential.py:791-828, 1078-1086``).  ``frame = (R[3, 3], t[3])``
    maps the object-local rays to the coordinates of the first surface.
    """
    (w, fx, fy, px, py), idx = cell_samples(vertices, begin, count, random, seed, chromatic)
    if at_infinity:
        x, y, (dx, dy, dz) = px, py, direction(fx, fy)
    else:
        x, y, (dx, dy, dz) = fx, fy, direction(px, py)
    z = np.zeros_like(x)
    intensity = np.ones_like(x)
    if weight_scene is not None:
        intensity = intensity * np.asarray(weight_scene, dtype=np.float64)[idx[0], idx[1], idx[2]]
    if weight_pupil is not None:
        wp = np.asarray(weight_pupil, dtype=np.float64)
        intensity = intensity * (wp[idx[0], idx[3], idx[4]] if wp.ndim == 3 else wp[idx[3], idx[4]])
    if frame is not None:
        r, t = np.asarray(frame[0], dtype=np.float64), np.asarray(frame[1], dtype=np.float64)
        x, y, z = (r[i, 0] * x + r[i, 1] * y + r[i, 2] * z + t[i] for i in range(3))
        dx, dy, dz = (r[i, 0] * dx + r[i, 1] * dy + r[i, 2] * dz for i in range(3))
    flat = lambda a: np.ascontiguousarray(a, dtype=np.float64).reshape(-1)  # noqa: E731
    n = flat(x).size
    return {
        "wavelength": flat(w),
        "px": flat(x),
        "py": flat(y),
        "pz": flat(z),
        "dx": flat(dx),
        "dy": flat(dy),
        "dz": flat(dz),
        "intensity": flat(intensity),
        "attenuation": np.zeros(n),
        "index_refraction": np.ones(n),
        "unvignetted": np.ones(n, dtype=bool),
    }


# -- cell areas (optika/vectors/_vectors_object.py:42-133) --------------------------------


def volume_cell_1d(v):
    """Signed length of every cell of a 1-D vertex array."""
    v = np.asarray(v, dtype=np.float64)
    return v[1:] - v[:-1]


def volume_cell_2d(x, y):
    """Signed area of every cell of a 2-D vertex grid ``x[i, j], y[i, j]`` (half the cross product of the diagonals)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    d1x, d1y = x[1:, 1:] - x[:-1, :-1], y[1:, 1:] - y[:-1, :-1]
    d2x, d2y = x[:-1, 1:] - x[1:, :-1], y[:-1, 1:] - y[1:, :-1]
    return 0.5 * (d1x * d2y - d1y * d2x)


def _solid_angle_triangle(a, b, c):
    """Van Oosterom & Strackee: signed solid angle of the spherical triangle of three unit vectors."""
    num = (
        a[0] * (b[1] * c[2] - b[2] * c[1])
        + a[1] * (b[2] * c[0] - b[0] * c[2])
        + a[2] * (b[0] * c[1] - b[1] * c[0])
    )
    dot = lambda p, q: p[0] * q[0] + p[1] * q[1] + p[2] * q[2]  # noqa: E731
    den = 1.0 + dot(a, b) + dot(b, c) + dot(c, a)
    return 2.0 * np.arctan2(num, den)


def solid_angle_cell(ax, ay):
    """Signed solid angle [sr] of every cell of a 2-D grid of angle vertices ``ax[i, j], ay[i, j]`` [rad]."""
    d = np.stack(direction(np.asarray(ax, dtype=np.float64), np.asarray(ay, dtype=np.float64)))
    v00, v10, v11, v01 = d[:, :-1, :-1], d[:, 1:, :-1], d[:, 1:, 1:], d[:, :-1, 1:]
    return _solid_angle_triangle(v00, v10, v11) + _solid_angle_triangle(v00, v11, v01)


def cell_area(vertices, field_is_angular: bool, pupil_is_angular: bool):
    """
    ``(area_wavelength[n0], area_field[n1, n2], area_pupil[n3, n4])`` of a separable grid:
    the three factors of ``ObjectVectorArray.cell_area`` (``_vectors_object.py:98-133``).
    """
    w, fx, fy, px, py = [np.asarray(v, dtype=np.float64) for v in vertices]
    FX, FY = (fx, fy) if fx.ndim == 2 else np.meshgrid(fx, fy, indexing="ij")
    PX, PY = (px, py) if px.ndim == 2 else np.meshgrid(px, py, indexing="ij")
    area_field = solid_angle_cell(FX, FY) if field_is_angular else volume_cell_2d(FX, FY)
    area_pupil = solid_angle_cell(PX, PY) if pupil_is_angular else volume_cell_2d(PX, PY)
    return volume_cell_1d(w), np.abs(area_field), np.abs(area_pupil)
