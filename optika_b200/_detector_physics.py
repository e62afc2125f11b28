"""
Detector physics after binning: what a back-illuminated CCD does with the
photons ``AbstractImagingSensor.collect`` counted.

Mirrors, for the step right behind the hot path,

* ``optika.sensors.charge_diffusion`` / ``mean_charge_capture`` / ``kernel_diffusion``
  (``optika/sensors/materials/_diffusion.py:13-138, 141-264, 316-418``): closed forms, host;
* the silicon pair-creation model of Ramanathan & Kurinsky 2020
  (``optika/sensors/materials/_ramanathan_2020/_ramanathan_2020.py:87-466``): bandgap, pair-creation
  energy, ideal quantum yield, Fano factor and the probability of n pairs, interpolated from the
  paper's tables (``optika_b200/data/ramanathan_2020/p*.dat``, the reference's data files);
* ``electrons_measured`` (``:468-690``) whose numba kernel (``:762-876``) runs on the device as
  ``optk_electrons_measured`` (``csrc/electrons.cu``): one thread per pixel, counter-based random
  numbers.

Units follow the rest of the package: lengths in mm (wavelengths too), energies in eV, temperatures in K.
"""

from __future__ import annotations
import ctypes as C
import functools
import pathlib
import numpy as np
from . import named as na
from . import units as u
from . import _lib as L

__all__ = [
    "charge_diffusion", "mean_charge_capture", "kernel_diffusion",
    "energy_bandgap", "energy_pair", "energy_pair_inf", "quantum_yield_ideal", "fano_factor", "fano_factor_inf",
    "probability_of_n_pairs", "electrons_measured", "photon_energy",
]

_HC_EV_MM = 1.2398419843320026e-3  # h c in eV mm


def photon_energy(wavelength):
    """Photon energy in eV of a vacuum wavelength in mm (astropy's ``u.spectral()`` equivalency)."""
    return _HC_EV_MM / na.as_named_array(u.length(wavelength))


# ---------------------------------------------------------------------------
# charge diffusion, closed forms
# ---------------------------------------------------------------------------
def charge_diffusion(absorption, thickness_substrate, thickness_depletion):
    """
    Standard deviation (mm) of the charge-diffusion kernel of a back-illuminated CCD, averaged over the
    absorption depth (``_diffusion.py:13-138``): ``sqrt(f (a f + exp(-a f) - 1) / (a (1 - exp(-a s))))`` with
    the field-free thickness ``f = s - depletion``.  `absorption` in 1 / mm.
    """
    s = u.length(thickness_substrate)
    f = s - u.length(thickness_depletion)
    a = absorption
    return np.sqrt(f * (a * f + np.exp(-a * f) - 1) / (a * (1 - np.exp(-a * s))))


def mean_charge_capture(width_diffusion, width_pixel):
    """Fraction of the charge of a photon event kept by the central pixel (``_diffusion.py:141-264``)."""
    from scipy.special import erf

    a = na.as_named_array(u.length(width_pixel) / u.length(width_diffusion))
    t1 = np.sqrt(2 / np.pi) * (np.exp(-np.square(a) / 2) - 1) / a
    t2 = na.ScalarArray(erf(a.ndarray / np.sqrt(2)), a.axes)
    return np.square(t1 + t2)


def _kernel_1d(width_diffusion, width_pixel, index_pixel):
    """``_diffusion.py:267-313``: Gaussian convolved with a pixel, integrated over the pixels `index_pixel`."""
    from scipy.special import erf

    x = na.as_named_array(u.length(width_pixel) / u.length(width_diffusion))
    n = na.as_named_array(index_pixel)
    x2 = np.square(x)
    c = 1 / (x * np.sqrt(2 * np.pi))

    def g(m):
        return np.exp(-x2 * m / 2)

    def e(m):
        arg = na.as_named_array(x * m / np.sqrt(2))
        return m * na.ScalarArray(erf(arg.ndarray), arg.axes)

    return c * (g(np.square(n - 1)) - 2 * g(np.square(n)) + g(np.square(n + 1))) + e(n - 1) / 2 - e(n) + e(n + 1) / 2


def kernel_diffusion(width_diffusion, width_pixel, axis_x: str, axis_y: str) -> na.FunctionArray:
    """The 3 x 3 charge-diffusion kernel (``_diffusion.py:316-418``): pixel offsets in, weights out."""
    index_x = na.linspace(-1, 1, axis=axis_x, num=3)
    index_y = na.linspace(-1, 1, axis=axis_y, num=3)
    kx = _kernel_1d(width_diffusion, width_pixel, index_x)
    ky = _kernel_1d(width_diffusion, width_pixel, index_y)
    return na.FunctionArray(inputs=na.Cartesian2dVectorArray(index_x, index_y), outputs=kx * ky)


# ---------------------------------------------------------------------------
# Ramanathan & Kurinsky 2020
# ---------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _tables():
    """(energy[eV], n[20], temperature[3], probability[temperature, energy, n]) from the paper's files (:27-84)."""
    directory = pathlib.Path(__file__).parent / "data" / "ramanathan_2020"
    files = [np.loadtxt(directory / name) for name in ("p0K.dat", "p100K.dat", "p300K.dat")]
    energy = files[0][:, 0]
    probability = np.stack([f[:, 1:] for f in files])
    n = np.arange(1, probability.shape[-1] + 1)
    return energy, n, np.array([0.0, 100.0, 300.0]), probability


def energy_bandgap(temperature=300.0):
    """Bandgap of silicon in eV (``:87-144``): ``1.1692 - 4.9e-4 T^2 / (T + 655)``."""
    T = np.asarray(temperature, dtype=float)
    return 1.1692 - 4.9e-4 * np.square(T) / (T + 655.0)


def energy_pair_inf(temperature=300.0):
    """Asymptotic pair-creation energy in eV (``:220-253``): ``1.7 E_g + 0.084 A + 1.3`` with A = 5.2 eV^2."""
    return 1.7 * energy_bandgap(temperature) + 0.084 * 5.2 + 1.3


def fano_factor_inf(temperature=300.0):
    """Asymptotic Fano factor (``:383-418``): ``-0.028 E_g + 0.0015 A + 0.14``."""
    return -0.028 * energy_bandgap(temperature) + 0.0015 * 5.2 + 0.14


def _interp_temperature(values, temperature):
    """Linear interpolation of ``values[temperature, ...]`` at scalar `temperature` (ends clamped, like na.interp)."""
    _, _, t, _ = _tables()
    T = float(np.clip(temperature, t[0], t[-1]))
    k = int(np.clip(np.searchsorted(t, T, side="right") - 1, 0, len(t) - 2))
    w = (T - t[k]) / (t[k + 1] - t[k])
    return (1 - w) * values[k] + w * values[k + 1]


def energy_pair(wavelength, temperature=300.0):
    """Mean pair-creation energy in eV (``:147-217``): tabulated energy / <n> below the tables' end, else the asymptote."""
    energy = photon_energy(wavelength)
    e, n, _, p = _tables()
    iqy = (n * p).sum(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        table = _interp_temperature(e / iqy, temperature)
    ok = np.isfinite(table)
    value = np.interp(energy.ndarray, e[ok], table[ok], right=float(energy_pair_inf(temperature)))
    return na.ScalarArray(value, energy.axes)


def quantum_yield_ideal(wavelength, temperature=300.0):
    """Electrons per absorbed photon (``:256-306``): photon energy / pair-creation energy."""
    return photon_energy(wavelength) / energy_pair(wavelength, temperature)


def fano_factor(wavelength, temperature=300.0):
    """Fano factor (``:309-380``): (<n^2> - <n>^2) / <n> of the tabulated distribution, asymptote beyond it."""
    energy = photon_energy(wavelength)
    e, n, t, p = _tables()
    iqy = (n * p).sum(-1)
    v = (np.square(n) * p).sum(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = (v - np.square(iqy)) / iqy
    per_temperature = []
    for k in range(len(t)):
        ok = np.isfinite(f[k])
        per_temperature.append(np.interp(energy.ndarray, e[ok], f[k][ok], right=float(fano_factor_inf(t[k]))))
    return na.ScalarArray(_interp_temperature(np.stack(per_temperature), temperature), energy.axes)


def probability_of_n_pairs(wavelength, temperature=300.0):
    """``(n[20], probability[..., 20])`` of creating n pairs (``:421-465``): tables interpolated in temperature, then energy."""
    energy = photon_energy(wavelength)
    e, n, _, p = _tables()
    at_t = _interp_temperature(p, temperature)  # [energy, n]
    flat = energy.ndarray.reshape(-1)
    out = np.stack([np.interp(flat, e, at_t[:, k]) for k in range(len(n))], axis=-1)
    return n, out.reshape(energy.ndarray.shape + (len(n),))


# ---------------------------------------------------------------------------
# the Monte-Carlo electron kernel on the device
# ---------------------------------------------------------------------------
def _per_plane(value, planes_shape: dict, what: str) -> np.ndarray:
    """A parameter broadcast over the image planes (everything but the two pixel axes), flattened."""
    v = na.as_named_array(value)
    extra = [ax for ax in v.axes if ax not in planes_shape]
    if extra:
        raise NotImplementedError(
            f"{what} varies along the pixel axes {extra}: the device kernel takes one value per image plane"
        )
    dims = tuple(planes_shape.values())
    return np.broadcast_to(na.aligned(v, planes_shape), dims).astype(float).reshape(-1)


def electrons_measured(
    photons_absorbed,
    wavelength,
    absorption=None,
    thickness_implant=40 * u.nm,
    thickness_depletion=None,
    thickness_substrate=7 * u.um,
    width_pixel=0.0,
    cce_backsurface=1.0,
    temperature=300.0,
    axis_xy: None | tuple = None,
    shape_random: None | dict = None,
    wrap: bool = False,
    seed: int = 0,
    device=None,
):
    """
    Monte-Carlo number of electrons measured by a back-illuminated CCD for the photons absorbed in every
    pixel (``optika/sensors/materials/_ramanathan_2020/_ramanathan_2020.py:468-690``; the numba kernel
    ``:762-876`` runs as ``optk_electrons_measured``).

    `photons_absorbed`: named integer array; `axis_xy` names its two pixel axes (without it every element
    is its own 1 x 1 sensor, as in the reference); `wavelength`, `absorption` (default: silicon,
    ``4 pi k / lambda``), the thicknesses, `width_pixel` (scalar or 2-D vector), `cce_backsurface` and
    `temperature` may vary along every OTHER axis (one value per image plane).  `shape_random` adds axes of
    independent realisations.  `seed` selects the counter-based random stream (the reference draws from
    Python's global generator).  Returns a named array of electron counts (float64, like the reference).
    """
    from . import _engine
    from .chemicals import Chemical

    torch = _engine._torch()
    device = _engine.require_cuda(device)
    photons = na.as_named_array(photons_absorbed)
    w = na.as_named_array(u.length(wavelength))
    if absorption is None:
        k = np.imag(Chemical("Si").n(w).ndarray)
        absorption = na.ScalarArray(4 * np.pi * k / w.ndarray, w.axes)  # Chemical.absorption
    if thickness_depletion is None:
        thickness_depletion = thickness_substrate  # :601-602
    if isinstance(width_pixel, na.Cartesian2dVectorArray):
        wp_x, wp_y = width_pixel.x, width_pixel.y
    else:
        wp_x = wp_y = width_pixel
    parameters = dict(
        wavelength=w, absorption=absorption, thickness_implant=u.length(thickness_implant),
        thickness_depletion=u.length(thickness_depletion), thickness_substrate=u.length(thickness_substrate),
        width_pixel_x=u.length(wp_x), width_pixel_y=u.length(wp_y), cce_backsurface=cce_backsurface,
    )
    full = na.broadcast_shapes(photons.shape, *[na.shape(v) for v in parameters.values()], shape_random or {})
    if axis_xy is not None:
        axis_x, axis_y = axis_xy
        full[axis_x] = full.pop(axis_x)  # the pixel axes last, as in the reference (:627-630)
        full[axis_y] = full.pop(axis_y)
        n_x, n_y = full[axis_x], full[axis_y]
        planes_shape = {ax: n for ax, n in full.items() if ax not in (axis_x, axis_y)}
    else:
        n_x = n_y = 1
        planes_shape = dict(full)
    n_plane = int(np.prod(list(planes_shape.values()), dtype=np.int64)) if planes_shape else 1
    values = {name: _per_plane(v, planes_shape, name) for name, v in parameters.items()}
    T = float(np.asarray(temperature, dtype=float))
    energy = _HC_EV_MM / values["wavelength"]
    n_values, pmf = probability_of_n_pairs(na.ScalarArray(values["wavelength"], "_plane"), T)
    cmf = np.cumsum(pmf, axis=-1)
    pair_inf = float(energy_pair_inf(T))  # 1 keV is beyond the tables: the asymptote (:663-665)
    fano_inf = float(fano_factor(_HC_EV_MM / 1000.0, T).ndarray)

    dims = tuple(full.values())
    counts = np.broadcast_to(na.aligned(photons, full), dims).astype(np.int64).reshape(n_plane, n_x, n_y)
    with _engine.device_guard(device):
        photons_dev = torch.from_numpy(np.ascontiguousarray(counts)).to(device)
        electrons = torch.zeros((n_plane, n_x, n_y), dtype=torch.int64, device=device)
        cmf_dev = torch.from_numpy(np.ascontiguousarray(cmf, dtype=np.float64)).to(device)
        n_dev = torch.from_numpy(np.ascontiguousarray(n_values, dtype=np.float64)).to(device)
        planes = (L.CcdPlane * n_plane)()
        for i in range(n_plane):
            P = planes[i]
            P.energy, P.absorption = float(energy[i]), float(values["absorption"][i])
            P.thickness_implant = float(values["thickness_implant"][i])
            P.thickness_depletion = float(values["thickness_depletion"][i])
            P.thickness_substrate = float(values["thickness_substrate"][i])
            P.width_pixel_x, P.width_pixel_y = float(values["width_pixel_x"][i]), float(values["width_pixel_y"][i])
            P.cce_backsurface = float(values["cce_backsurface"][i])
            P.energy_pair_inf, P.fano_inf = pair_inf, fano_inf
            P.n_pmf = cmf.shape[-1]
            P.cmf = cmf_dev.data_ptr() + 8 * i * cmf.shape[-1]
            P.n_values = n_dev.data_ptr()
        L.check(
            L.lib().optk_electrons_measured(
                n_plane, n_x, n_y, planes, photons_dev.data_ptr(), electrons.data_ptr(), 1 if wrap else 0,
                int(seed) & 0xFFFFFFFFFFFFFFFF, _engine._stream_ptr(device),
            )
        )
        result = electrons.cpu().numpy().astype(np.float64).reshape(dims)
    return na.ScalarArray(result, tuple(full))
