"""
ORACLE (test infrastructure, not product code): NumPy restatement of the
reference's detector binning, ``AbstractImagingSensor.collect``,
``optika/sensors/_sensors.py:92-171``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module; the product (``optika_b200``) never does.

``na.histogram`` is third-party (named_arrays ~= 2.1, source not in
/root/reference); it is restated as ``numpy.histogramdd`` semantics: bins
``[e_i, e_{i+1})`` with the last bin closed, samples outside the edges or NaN
dropped.  Bin edges are exactly ``linspace(bound_lower, bound_upper, num_pixel + 1)``
(``_sensors.py:141-149``), so bin lookup compares against those edge VALUES.
The reference's tests check only types for this step
(``optika/sensors/_sensors_test.py:54-94``): parity on bin edges is UNPINNED and
tests enumerate the rays within tolerance of an edge via :func:`bin_margin`.
"""

from __future__ import annotations
import numpy as np

__all__ = ["pixel_edges", "collect", "counts", "bin_margin"]


def pixel_edges(half_width: float, num_pixel: int) -> np.ndarray:
    """``linspace(bound_lower, bound_upper, num_pixel + 1)`` with ``bound = -/+ width_pixel * num_pixel / 2``
    (``_sensors.py:83-90, 141-149``)."""
    return np.linspace(-half_width, half_width, num_pixel + 1)


def _sample(rays: dict, where=True):
    where = np.asarray(where) & rays["unvignetted"]  # _sensors.py:125
    # IdealSensorMaterial.direction_refracted = -direction . normal, normal = (0, 0, -1)
    # (optika/sensors/materials/_materials.py:1603-1616; sag is NoSag, _sensors.py:45-46)
    cos_refracted = rays["dz"] + 0j
    flux = rays["intensity"] * where  # _sensors.py:139
    sample = np.stack(
        [rays["wavelength"].ravel(), rays["px"].ravel(), rays["py"].ravel()], axis=-1
    )
    return sample, flux.ravel(), cos_refracted.ravel()


def collect(rays: dict, edges_wavelength, edges_x, edges_y, where=True):
    """
    ``collect``: returns ``(image, direction)`` with `image` of shape
    ``(n_wavelength, n_x, n_y)`` and the flux-weighted mean refracted cosine
    (complex), ``_sensors.py:151-171``; plus the three raw histogram planes.
    """
    sample, flux, cosr = _sample(rays, where)
    bins = [np.asarray(edges_wavelength), np.asarray(edges_x), np.asarray(edges_y)]
    with np.errstate(invalid="ignore"):
        good = np.all(np.isfinite(sample), axis=-1)
    s = sample[good]
    image, _ = np.histogramdd(s, bins=bins, weights=flux[good])
    moment_real, _ = np.histogramdd(s, bins=bins, weights=(flux * np.real(cosr))[good])
    moment_imag, _ = np.histogramdd(s, bins=bins, weights=(flux * np.imag(cosr))[good])
    nonempty = image > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        direction_real = np.where(nonempty, moment_real / image, 1)
        direction_imag = np.where(nonempty, moment_imag / image, 0)
    direction = direction_real + direction_imag * 1j
    return image, direction, (image, moment_real, moment_imag)


def counts(rays: dict, edges_wavelength, edges_x, edges_y, where=True) -> np.ndarray:
    """Integer number of unvignetted rays per bin (the bit-exact quantity of the north star)."""
    sample, flux, _ = _sample(rays, where)
    where = (np.asarray(where) & rays["unvignetted"]).ravel()
    with np.errstate(invalid="ignore"):
        good = np.all(np.isfinite(sample), axis=-1) & where
    h, _ = np.histogramdd(
        sample[good],
        bins=[np.asarray(edges_wavelength), np.asarray(edges_x), np.asarray(edges_y)],
    )
    return h.astype(np.int64)


def bin_margin(values, edges) -> np.ndarray:
    """Distance from each value to the nearest bin edge (test infrastructure)."""
    values = np.asarray(values, dtype=float)
    edges = np.asarray(edges, dtype=float)
    i = np.clip(np.searchsorted(edges, values), 1, len(edges) - 1)
    return np.minimum(np.abs(values - edges[i - 1]), np.abs(values - edges[i]))
