"""
Minimal unit constants for the host side of the engine.

The reference expresses every length, wavelength and angle as an
``astropy.units.Quantity`` (e.g. ``optika/sags/_spherical.py:113-137`` converts
to the unit of the radius before calling numexpr).  ``astropy`` is not available
where this engine is built, and the device works in exactly one unit system, so
the host side uses plain floats in **engine units**:

* lengths (positions, radii, wavelengths, ruling spacings, thicknesses): millimetres
* angles: radians
* attenuation: 1 / mm

``500 * u.nm`` therefore evaluates to the float ``5e-4``.  Real astropy
quantities are accepted wherever a length or angle is expected and converted by
:func:`length` / :func:`angle`.
"""

from __future__ import annotations
import math
import numpy as np

__all__ = [
    "m", "cm", "mm", "um", "nm", "AA", "angstrom",
    "rad", "deg", "arcmin", "arcsec",
    "dimensionless_unscaled", "s", "photon", "electron",
    "length", "angle",
]

# lengths, in millimetres
m = 1e3
cm = 10.0
mm = 1.0
um = 1e-3
nm = 1e-6
AA = 1e-7
angstrom = AA

# angles, in radians
rad = 1.0
deg = math.pi / 180
arcmin = deg / 60
arcsec = deg / 3600

dimensionless_unscaled = 1.0
s = 1.0
photon = 1.0
electron = 1.0


def _is_quantity(a) -> bool:
    return hasattr(a, "unit") and hasattr(a, "to_value")


def length(a):
    """Convert `a` to engine length units (mm). Astropy quantities are converted."""
    if _is_quantity(a):
        import astropy.units as au  # pragma: no cover

        return a.to_value(au.mm)  # pragma: no cover
    return a


def angle(a):
    """Convert `a` to engine angle units (radians). Astropy quantities are converted."""
    if _is_quantity(a):
        import astropy.units as au  # pragma: no cover

        return a.to_value(au.rad)  # pragma: no cover
    return a


def asfloat(a) -> np.ndarray:
    return np.asarray(a, dtype=np.float64)
