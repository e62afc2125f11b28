"""
ORACLE (test infrastructure, not product code): NumPy complex128 restatement of
the reference's transfer-matrix multilayer model.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module; the product (``optika_b200``) never does.

Follows (paths relative to ``/root/reference``):
``optika/materials/_multilayers.py:187-237`` (``multilayer_coefficients``),
``:485-532`` (``multilayer_efficiency``),
``optika/materials/_layers.py:229-277`` (``Layer.transfer``), ``:473-499``
(``LayerSequence.transfer``), ``:611-645`` (``PeriodicLayerSequence.transfer``),
``optika/materials/matrices.py:127-161`` (``refraction``), ``:241-246``
(``propagation``), ``optika/materials/profiles.py:103-126`` and the four
``_derivative_fourier_transform`` bodies, ``optika/materials/_snells_law.py:13-38``
(``snells_law_scalar``), ``optika/chemicals/_chemicals.py:101-144``.

Layers are plain tuples ``(n, thickness, profile_kind, profile_width)`` where `n`
is a complex array broadcastable against the evaluation grid (already
interpolated from the ``.nk`` table), `thickness` a float array in the same
length unit as `wavelength`, and `profile_kind` one of 0 (none), 1 (erf),
2 (exponential), 3 (linear), 4 (sinusoidal).

Pinning: the four IMD golden tables the reference tests against
(``optika/materials/_tests/test_multilayers.py:178-287``, ``rtol=1e-4``) are
committed as ``tests/golden/imd_*.npz`` and checked in
``tests/test_oracle_multilayer.py``; periodic == explicit
(``optika/materials/_tests/test_layers.py:240-291``) likewise.
``Cartesian2dMatrixArray.power`` is third-party (named_arrays ~= 2.1): restated as
the n-fold product, which is what the reference's own test pins it to.
"""

from __future__ import annotations
import numpy as np

__all__ = [
    "snells_law_scalar",
    "interface_reflectivity",
    "refraction",
    "propagation",
    "layer_transfer",
    "sequence_transfer",
    "periodic_transfer",
    "multilayer_coefficients",
    "multilayer_efficiency",
    "load_nk",
    "interp_nk",
]


def snells_law_scalar(cos_incidence, index_refraction, index_refraction_new):
    """``optika/materials/_snells_law.py:13-38`` (``np.emath.sqrt``: complex for negative args)."""
    cos_incidence = np.asarray(cos_incidence)
    sin_incidence = np.emath.sqrt(1 - np.square(cos_incidence))
    sin_transmitted = index_refraction * sin_incidence / index_refraction_new
    cos_transmitted = np.emath.sqrt(1 - np.square(sin_transmitted))
    return cos_transmitted


def interface_reflectivity(kind: int, width, wavelength, direction, n):
    """
    ``AbstractInterfaceProfile.reflectivity``, ``optika/materials/profiles.py:103-126``:
    ``k = -2 pi n direction / wavelength``, ``s = Re(-2 k)``, then the profile's
    ``_derivative_fourier_transform(s)`` (``:221-222``, ``:320-325``, ``:424-431``, ``:532-541``).
    """
    k = -2 * np.pi * n * direction / wavelength
    s = np.real(-2 * k)
    with np.errstate(invalid="ignore", divide="ignore"):
        if kind == 1:
            return np.exp(-np.square(s * width) / 2)
        if kind == 2:
            return 1 / (1 + np.square(s * width) / 2)
        if kind == 3:
            x = np.sqrt(3) * width * s
            return np.sin(x) / x
        if kind == 4:
            a = np.pi / (np.square(np.pi) - 8)
            x = a * width * s
            x1 = x - np.pi / 2
            x2 = x + np.pi / 2
            return np.pi * (np.sin(x1) / x1 + np.sin(x2) / x2) / 4
    raise ValueError(f"unknown profile kind {kind}")


def _matmul(a, b):
    """2x2 complex matrix product on tuples ((xx, xy), (yx, yy)) of arrays."""
    (a00, a01), (a10, a11) = a
    (b00, b01), (b10, b11) = b
    return (
        (a00 * b00 + a01 * b10, a00 * b01 + a01 * b11),
        (a10 * b00 + a11 * b10, a10 * b01 + a11 * b11),
    )


def _where(cond, a, b):
    return tuple(tuple(np.where(cond, x, y) for x, y in zip(ra, rb)) for ra, rb in zip(a, b))


_IDENTITY = ((1.0 + 0j, 0.0 + 0j), (0.0 + 0j, 1.0 + 0j))


def refraction(wavelength, direction_left, direction_right, polarized_s, n_left, n_right,
               profile_kind=0, profile_width=0.0):
    """``optika/materials/matrices.py:127-161``."""
    direction_i = np.where(polarized_s, direction_left, np.conj(direction_left))
    direction_j = np.where(polarized_s, direction_right, np.conj(direction_right))
    n_i = n_left
    n_j = n_right
    impedance_i = np.where(polarized_s, n_i, 1 / np.asarray(n_i, dtype=complex))
    impedance_j = np.where(polarized_s, n_j, 1 / np.asarray(n_j, dtype=complex))
    q_i = direction_i * impedance_i
    q_j = direction_j * impedance_j
    with np.errstate(invalid="ignore", divide="ignore"):
        a_ij = q_i + q_j
        r_ij = (q_i - q_j) / a_ij
        t_ij = 2 * q_i / a_ij
        if profile_kind:
            r_ij = r_ij * interface_reflectivity(
                profile_kind, profile_width, wavelength, direction_i, n_i
            )
        one = np.ones_like(r_ij)
        return ((one / t_ij, r_ij / t_ij), (r_ij / t_ij, one / t_ij))


def propagation(wavelength, direction, thickness, n):
    """``optika/materials/matrices.py:241-246``."""
    with np.errstate(over="ignore", invalid="ignore"):
        beta = 2 * np.pi * thickness * n * direction / wavelength
        zero = np.zeros_like(beta)
        return ((np.exp(-1j * beta), zero), (zero, np.exp(+1j * beta)))


def layer_transfer(layer, wavelength, direction, polarized_s, n, where=True):
    """``Layer.transfer``, ``optika/materials/_layers.py:229-277``."""
    n_internal, thickness, kind, width = layer
    direction_internal = snells_law_scalar(direction, n, n_internal)
    w = refraction(wavelength, direction, direction_internal, polarized_s, n, n_internal, kind, width)
    w = _where(where, w, _IDENTITY)
    u = propagation(wavelength, direction_internal, thickness, n_internal)
    with np.errstate(invalid="ignore"):
        where_propagation = np.abs(u[0][0]) < 1e10
    where = where & where_propagation
    with np.errstate(invalid="ignore", over="ignore"):
        transfer = _matmul(w, u)
    transfer = _where(where, transfer, w)
    return n_internal, direction_internal, transfer, where


def sequence_transfer(layers, wavelength, direction, polarized_s, n, where=True):
    """``LayerSequence.transfer``, ``optika/materials/_layers.py:473-499``."""
    result = _IDENTITY
    for layer in layers:
        n, direction, m, where = layer_transfer(layer, wavelength, direction, polarized_s, n, where)
        with np.errstate(invalid="ignore", over="ignore"):
            result = _matmul(result, m)
    return n, direction, result, where


def periodic_transfer(layers, num_periods, wavelength, direction, polarized_s, n, where=True):
    """``PeriodicLayerSequence.transfer``, ``optika/materials/_layers.py:611-645``."""
    n, direction, start, where = sequence_transfer(layers, wavelength, direction, polarized_s, n, where)
    n, direction, periodic, where = sequence_transfer(layers, wavelength, direction, polarized_s, n, where)
    power = _IDENTITY
    with np.errstate(invalid="ignore", over="ignore"):
        for _ in range(num_periods - 1):
            power = _matmul(power, periodic)
        return n, direction, _matmul(start, power), where


def _stack_transfer(stack, wavelength, direction, polarized_s, n, where=True):
    """
    `stack` is a list whose items are either a layer tuple or
    ``("periodic", [layers...], num_periods)``.
    """
    result = _IDENTITY
    for item in stack:
        if isinstance(item[0], str) and item[0] == "periodic":
            n, direction, m, where = periodic_transfer(
                item[1], item[2], wavelength, direction, polarized_s, n, where
            )
        else:
            n, direction, m, where = layer_transfer(item, wavelength, direction, polarized_s, n, where)
        with np.errstate(invalid="ignore", over="ignore"):
            result = _matmul(result, m)
    return n, direction, result, where


def multilayer_coefficients(wavelength, direction, n, stack, substrate):
    """
    ``optika/materials/_multilayers.py:187-237``.  `substrate` is a layer tuple
    whose thickness is forced to 0 (``:190-193``); ``None`` means vacuum.
    Returns ``(r_s, r_p, t_s, t_p)``.
    """
    wavelength = np.asarray(wavelength, dtype=float)
    direction = np.asarray(direction)
    n = np.asarray(n)
    if substrate is None:
        substrate = (1.0 + 0j, 0.0, 0, 0.0)
    substrate = (substrate[0], 0.0, substrate[2], substrate[3])
    out = []
    for polarized_s in (True, False):
        n_, d_, m_layers, where = _stack_transfer(stack, wavelength, direction, polarized_s, n)
        n_, d_, m_substrate, where = layer_transfer(substrate, wavelength, d_, polarized_s, n_, where)
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            m = _matmul(m_layers, m_substrate)
            r = m[1][0] / m[0][0]
            t = 1 / m[0][0]
            t = np.where(where, t, 0)
        out.append((r, t))
    return out[0][0], out[1][0], out[0][1], out[1][1]


def multilayer_efficiency(wavelength, direction=1.0, n=1.0, stack=(), substrate=None):
    """
    ``optika/materials/_multilayers.py:485-532``.
    Returns ``(R_s, R_p, T_s, T_p)`` broadcast over the evaluation grid.
    """
    wavelength = np.asarray(wavelength, dtype=float)
    direction = np.asarray(direction)
    n = np.asarray(n)
    r_s, r_p, t_s, t_p = multilayer_coefficients(wavelength, direction, n, stack, substrate)
    n_substrate = 1.0 + 0j if substrate is None else substrate[0]
    direction_substrate = snells_law_scalar(direction, n, n_substrate)
    with np.errstate(invalid="ignore", divide="ignore"):
        q_ambient_s = direction * n
        q_ambient_p = np.conj(direction) * (1 / np.asarray(n, dtype=complex))
        q_substrate_s = direction_substrate * n_substrate
        q_substrate_p = np.conj(direction_substrate) * (1 / np.asarray(n_substrate, dtype=complex))
        R_s = np.square(np.abs(r_s))
        R_p = np.square(np.abs(r_p))
        T_s = np.square(np.abs(t_s)) * np.real(q_substrate_s / q_ambient_s)
        T_p = np.square(np.abs(t_p)) * np.real(q_substrate_p / q_ambient_p)
    return R_s, R_p, T_s, T_p


def load_nk(file) -> tuple[np.ndarray, np.ndarray]:
    """Parse an IMD ``.nk`` table, ``optika/chemicals/_chemicals.py:116-134``: (wavelength [A], n + ik)."""
    skip = 0
    with open(file, "r") as f:
        for line in f:
            if line.startswith(";"):
                skip += 1
            else:
                break
    w, n, k = np.loadtxt(file, skiprows=skip, unpack=True)
    return w, n + 1j * k


def interp_nk(wavelength_angstrom, table) -> np.ndarray:
    """``na.interp`` of ``n + ik`` (``_chemicals.py:138-142``): linear, clamped ends."""
    return np.interp(wavelength_angstrom, table[0], table[1])
