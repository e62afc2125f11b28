"""``optika.direction`` / ``optika.angles`` (``optika/_util.py:41-97``), host side, on small axes."""

from __future__ import annotations
import numpy as np
from . import named as na

__all__ = ["direction", "angles", "shape"]


def shape(a) -> dict[str, int]:
    return na.shape(a)


def direction(angles: na.Cartesian2dVectorArray) -> na.Cartesian3dVectorArray:
    """Azimuth/elevation (radians) -> direction cosines, ``d = R_y(phi_x) R_x(phi_y) z`` (``_util.py:41-73``)."""
    return na.Cartesian3dVectorArray(
        x=-np.cos(angles.y) * np.sin(angles.x),
        y=-np.sin(angles.y),
        z=+np.cos(angles.y) * np.cos(angles.x),
    )


def angles(direction: na.Cartesian3dVectorArray) -> na.Cartesian2dVectorArray:
    """Inverse of :func:`direction` (radians), ``_util.py:76-97``."""
    return na.Cartesian2dVectorArray(
        x=-np.arctan2(direction.x, direction.z),
        y=-np.arcsin(direction.y / direction.length),
    )
