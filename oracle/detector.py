"""
TEST INFRASTRUCTURE -- never imported by the product.

CPU restatement of the detector physics that follows binning:

* ``charge_diffusion`` / ``mean_charge_capture`` / the 3 x 3 ``kernel_diffusion``
  (``optika/sensors/materials/_diffusion.py:13-138, 141-264, 267-418``): closed forms, NumPy;
* the Monte-Carlo electron kernel ``_electrons_measured_numba``
  (``optika/sensors/materials/_ramanathan_2020/_ramanathan_2020.py:762-876``), statement by statement,
  vectorised over the photons of a pixel.

The reference draws from Python's ``random`` module inside numba threads: its stream is not
reproducible by anyone.  The restatement therefore uses the counter-based generator the device
kernel uses (Philox4x32-10 keyed by seed, counter = pixel, photon, draw -- documented in
``include/optk.h``), which makes oracle and device comparable COUNT FOR COUNT; parity with the
reference itself is statistical and is pinned by the reference's own tests for this path
(``_ramanathan_2020_test.py:180-262``: the spread of the diffused charge equals ``charge_diffusion``
within 5 %, a wrapped grid keeps more charge than a dropping one).
"""

from __future__ import annotations
import numpy as np
from scipy.special import erf
from .grid import philox4x32_10

__all__ = [
    "charge_diffusion", "mean_charge_capture", "kernel_diffusion", "electrons_measured", "uniform53", "normal_pair",
]


def charge_diffusion(absorption, thickness_substrate, thickness_depletion):
    """``_diffusion.py:128-138``: sqrt(f (a f + exp(-a f) - 1) / (a (1 - exp(-a s))))."""
    s = thickness_substrate
    f = s - thickness_depletion
    a = absorption
    return np.sqrt(f * (a * f + np.exp(-a * f) - 1) / (a * (1 - np.exp(-a * s))))


def mean_charge_capture(width_diffusion, width_pixel):
    """``_diffusion.py:256-264``."""
    a = width_pixel / width_diffusion
    t1 = np.sqrt(2 / np.pi) * (np.exp(-np.square(a) / 2) - 1) / a
    return np.square(t1 + erf(a / np.sqrt(2)))


def _kernel_1d(width_diffusion, width_pixel, n):
    """``_diffusion.py:267-313``."""
    x = width_pixel / width_diffusion
    x2 = np.square(x)
    c = 1 / (x * np.sqrt(2 * np.pi))
    g = lambda m: np.exp(-x2 * m / 2)  # noqa: E731
    e = lambda m: m * erf(x * m / np.sqrt(2))  # noqa: E731
    return c * (g(np.square(n - 1)) - 2 * g(np.square(n)) + g(np.square(n + 1))) + e(n - 1) / 2 - e(n) + e(n + 1) / 2


def kernel_diffusion(width_diffusion, width_pixel):
    """``_diffusion.py:395-418``: outer product of two 1-D kernels on the pixel offsets -1, 0, 1."""
    n = np.array([-1.0, 0.0, 1.0])
    k = _kernel_1d(width_diffusion, width_pixel, n)
    return k[:, None] * k[None, :]


# ---------------------------------------------------------------------------
# random numbers: the device's stream (include/optk.h, "detector physics")
# ---------------------------------------------------------------------------
def _words(pixel, photon, draw, seed):
    pixel = np.asarray(pixel, dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return philox4x32_10((pixel & mask, pixel >> np.uint64(32), np.asarray(photon, dtype=np.uint64),
                          np.asarray(draw, dtype=np.uint64)), key)


def uniform53(hi, lo):
    """(0, 1): ((hi << 21 | lo >> 11) + 0.5) 2^-53 from two 32-bit words."""
    bits = (np.asarray(hi, dtype=np.uint64) << np.uint64(21)) | (np.asarray(lo, dtype=np.uint64) >> np.uint64(11))
    return (bits.astype(np.float64) + 0.5) * 2.0**-53


def normal_pair(u1, u2):
    """Box-Muller: two independent standard normals from two uniforms in (0, 1)."""
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)


def electrons_measured(photons, plane: dict, wrap: bool, seed: int, plane_index: int = 0) -> np.ndarray:
    """
    ``_electrons_measured_numba`` for one image plane: `photons[n_x, n_y]` integers; `plane` holds the
    per-plane scalars ``energy`` (eV), ``absorption`` (1 / mm), ``thickness_implant``,
    ``thickness_depletion``, ``thickness_substrate``, ``width_pixel_x``, ``width_pixel_y`` (mm),
    ``cce_backsurface``, ``p_n`` / ``n`` (pair-number pmf), ``energy_pair_inf`` (eV), ``fano_inf``.
    Pixel counter = ``(plane_index * n_x + x) * n_y + y``.  Returns electron counts ``[n_x, n_y]``.
    """
    photons = np.asarray(photons)
    num_x, num_y = photons.shape
    result = np.zeros((num_x, num_y), dtype=np.int64)
    a = float(plane["absorption"])
    W = float(plane["thickness_implant"])
    h_0 = float(plane["cce_backsurface"])
    cmf = np.cumsum(np.asarray(plane["p_n"], dtype=float))  # :803
    n_i = np.asarray(plane["n"], dtype=float)
    z_substrate = float(plane["thickness_substrate"])
    z_ff = z_substrate - float(plane["thickness_depletion"])  # :808
    wp_x, wp_y = float(plane["width_pixel_x"]), float(plane["width_pixel_y"])
    d = 1 / a if a > 0 else 0.0  # :814
    fraction_absorbed = 1 - np.exp(-a * z_substrate)  # :818
    mean_inf = float(plane["energy"]) / float(plane["energy_pair_inf"])  # :820
    std_inf = np.sqrt(float(plane["fano_inf"]) * mean_inf)
    low_energy = float(plane["energy"]) <= 50  # :823
    for x in range(num_x):
        for y in range(num_y):
            num_photon = int(photons[x, y])
            if num_photon <= 0:
                continue
            pixel = (plane_index * num_x + x) * num_y + y
            j = np.arange(num_photon)
            w0 = _words(pixel, j, 0, seed)  # photon-level draws, block 0: pair number
            w1 = _words(pixel, j, 1, seed)  # block 1: depth, position in the pixel
            if low_energy:  # :826-833
                x_ij = uniform53(w0[0], w0[1])
                k = np.minimum(np.searchsorted(cmf, x_ij, side="right"), len(cmf) - 1)  # first k with cmf[k] > x
                n_ij = n_i[k]
            else:  # :835-840
                z0, _ = normal_pair(uniform53(w0[0], w0[1]), uniform53(w0[2], w0[3]))
                n_ij = np.rint(mean_inf + std_inf * z0)
            n_ij = np.maximum(n_ij, 0).astype(np.int64)
            y_ij = uniform53(w1[0], w1[1])
            z_ij = -d * np.log(1 - y_ij * fraction_absorbed) if a > 0 else y_ij * z_substrate  # :842-848
            h_ij = np.where(z_ij < W, h_0 + (1 - h_0) * z_ij / W if W > 0 else 1.0, 1.0)  # :850-853
            u_ = (w1[2].astype(np.float64) + 0.5) * 2.0**-32 - 0.5  # :857-858
            v_ = (w1[3].astype(np.float64) + 0.5) * 2.0**-32 - 0.5
            diffuses = (z_ij < z_ff) & (wp_x > 0) & (wp_y > 0)  # :861
            with np.errstate(invalid="ignore"):
                w_ = np.where(diffuses, z_ff * np.sqrt(np.maximum(1 - z_ij / z_ff, 0.0)) if z_ff != 0 else 0.0, 0.0)
            for jj in range(num_photon):
                n_pairs = int(n_ij[jj])
                if n_pairs == 0:
                    continue
                e = np.arange(n_pairs)
                keep = np.ones(n_pairs, dtype=bool)
                if h_ij[jj] < 1:  # binomial(n, h) as n Bernoulli trials (:855)
                    ws = _words(pixel, jj, (np.uint64(1) << np.uint64(31)) + np.uint64(2) + e.astype(np.uint64), seed)
                    keep = uniform53(ws[0], ws[1]) < h_ij[jj]
                e = e[keep]
                if e.size == 0:
                    continue
                if diffuses[jj]:
                    we = _words(pixel, jj, 2 + e, seed)
                    zp, zq = normal_pair(uniform53(we[0], we[1]), uniform53(we[2], we[3]))
                    p = np.rint(u_[jj] + (w_[jj] / wp_x) * zp).astype(np.int64)  # :864-868
                    q = np.rint(v_[jj] + (w_[jj] / wp_y) * zq).astype(np.int64)
                else:
                    p = q = np.zeros(e.size, dtype=np.int64)
                x_e, y_e = x + p, y + q
                if wrap:  # :876-878
                    np.add.at(result, (x_e % num_x, y_e % num_y), 1)
                else:
                    ok = (0 <= x_e) & (x_e < num_x) & (0 <= y_e) & (y_e < num_y)
                    np.add.at(result, (x_e[ok], y_e[ok]), 1)
    return result
