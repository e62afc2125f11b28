"""
Parity helpers shared by the GPU tests, ``smoke()`` and ``bench.py``
(not a test module).  Compares device rays with the NumPy oracle using the
north-star rules: positions and directions within 1e-9 relative in fp64; masks
bit-exact except for rays within tolerance of an aperture edge, which are
enumerated.
"""

from __future__ import annotations
import numpy as np
from oracle import raytrace as ora

RTOL = 1e-9  # BASELINE.json north_star: "within 1e-9 relative in fp64"


def _scale(a: np.ndarray) -> float:
    a = a[np.isfinite(a)]
    return float(np.max(np.abs(a))) if a.size else 1.0


def oracle_accumulate(surfaces, rays0: dict, converge: bool = True, local_last=None) -> dict:
    """Oracle states after every surface; optionally the last one in the sensor-local frame."""
    acc = ora.accumulate_rays(surfaces, rays0, converge=converge)
    return acc


def edge_rays(surfaces, states: dict, tol_rel: float = RTOL) -> np.ndarray:
    """
    Boolean array [n_surface, n_rays]: True where a ray is within tolerance of the
    edge of that surface's aperture (evaluated in the surface-local frame on the
    outgoing ray, as ``optika/surfaces.py:192-193`` does).  These rays are allowed
    to differ in ``unvignetted`` and are enumerated by the tests.
    """
    n_surf = len(surfaces)
    near = np.zeros(states["px"].shape, dtype=bool)
    for s, surface in enumerate(surfaces):
        ap = surface.aperture
        if ap is None:
            continue
        st = {k: v[s] for k, v in states.items()}
        local = ora._rays_transform(surface.transformation, st, inverse=True)
        if ora.is_angular(ap):
            v = (local["dx"], local["dy"], local["dz"])
            scale = 1.0
        else:
            v = (local["px"], local["py"], local["pz"])
            scale = max(_scale(local["px"]), _scale(local["py"]), 1e-300)
        with np.errstate(invalid="ignore"):
            margin = ora.aperture_margin(ap, *v)
            near[s] = ~(margin > tol_rel * scale)  # NaN margins count as "near"
    assert near.shape[0] == n_surf
    return near


def compare_states(device: dict, oracle: dict, surfaces=None, rtol: float = RTOL) -> dict:
    """
    Compare dicts of arrays shaped [n_surface, n_rays] (or [n_rays]).  Returns a
    report; raises AssertionError on a parity failure.
    """
    report = {}
    squeeze = oracle["px"].ndim == 1
    if squeeze:
        device = {k: v[None] for k, v in device.items()}
        oracle = {k: v[None] for k, v in oracle.items()}
    for group, names in (
        ("position", ("px", "py", "pz")),
        ("direction", ("dx", "dy", "dz")),
        ("wavelength", ("wavelength",)),
        ("intensity", ("intensity",)),
        ("attenuation", ("attenuation",)),
        ("index_refraction", ("index_refraction",)),
    ):
        worst = 0.0
        for s in range(oracle["px"].shape[0]):
            scale = max(_scale(np.stack([oracle[n][s] for n in names])), 1e-300)
            for n in names:
                a, b = device[n][s], oracle[n][s]
                nan_a, nan_b = ~np.isfinite(a), ~np.isfinite(b)
                assert np.array_equal(nan_a, nan_b), f"{n}: non-finite pattern differs at surface {s}"
                ok = ~nan_b
                err = np.max(np.abs(a[ok] - b[ok])) / scale if ok.any() else 0.0
                worst = max(worst, float(err))
                inf = np.isinf(b)
                assert np.array_equal(a[inf], b[inf]), f"{n}: infinities differ at surface {s}"
        report[group] = worst
        assert worst <= rtol, f"{group}: max relative error {worst:.3e} > {rtol:.1e}"
    mism = device["unvignetted"].astype(bool) != oracle["unvignetted"].astype(bool)
    report["mask_mismatches"] = int(mism.sum())
    if mism.any():
        assert surfaces is not None, "mask mismatch and no surfaces given to enumerate edge rays"
        near = edge_rays(surfaces, oracle)
        # a mismatch is excused only if the ray was near an edge at this or an earlier surface
        near_cum = np.logical_or.accumulate(near, axis=0)
        bad = mism & ~near_cum
        report["edge_rays"] = np.argwhere(mism).tolist()
        assert not bad.any(), f"{int(bad.sum())} mask mismatches away from any aperture edge"
    return report
