"""
Snell's law.

``snells_law`` keeps the signature of ``optika.materials.snells_law``
(``optika/materials/_snells_law.py:41-47``) and runs the vector form
(``:341-366``) on the device through a one-surface trace restricted to the
refraction stage.  ``snells_law_scalar`` (``:13-38``) is the complex scalar
form used by the multilayer model; it is evaluated inside
``optika_b200/csrc/multilayer.cu`` and exposed here only for small host-side
set-up computations.
"""

from __future__ import annotations
import numpy as np
from .. import named as na

__all__ = ["snells_law", "snells_law_scalar"]


def snells_law_scalar(cos_incidence, index_refraction, index_refraction_new):
    """Cosine of the refracted angle (``_snells_law.py:13-38``); host-side helper."""
    sin_incidence = np.emath.sqrt(1 - np.square(np.asarray(cos_incidence)))
    sin_transmitted = index_refraction * sin_incidence / index_refraction_new
    return np.emath.sqrt(1 - np.square(sin_transmitted))


def snells_law(
    direction: na.Cartesian3dVectorArray,
    index_refraction,
    index_refraction_new,
    normal: None | na.Cartesian3dVectorArray = None,
    is_mirror: bool = False,
) -> na.Cartesian3dVectorArray:
    """Vector form of Snell's law (``_snells_law.py:41-291``), evaluated on the device."""
    from .. import _engine

    return _engine.snells_law(direction, index_refraction, index_refraction_new, normal, is_mirror)
