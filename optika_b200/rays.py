"""
Ray bundles.

Mirrors ``optika.rays.RayVectorArray`` (``optika/rays/_ray_vectors.py:240-294``):
ten fp64 quantities plus the ``unvignetted`` mask, each a scalar or a named
array; the ray grid is their broadcast by axis name.  This is the host-side
description; :mod:`optika_b200._engine` flattens it into strided
structure-of-arrays buffers for the device without materialising broadcasts.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na

__all__ = ["RayVectorArray", "RayFunctionArray"]


@dataclasses.dataclass(eq=False)
class RayVectorArray:
    """An ensemble of light rays (fields as in ``_ray_vectors.py:256-278``)."""

    wavelength: float | na.ScalarArray = 0
    position: na.Cartesian3dVectorArray = dataclasses.field(
        default_factory=lambda: na.Cartesian3dVectorArray(0, 0, 0)
    )
    direction: na.Cartesian3dVectorArray = dataclasses.field(
        default_factory=lambda: na.Cartesian3dVectorArray(0, 0, 0)
    )
    intensity: float | na.ScalarArray = 1
    attenuation: float | na.ScalarArray = 0
    index_refraction: float | na.ScalarArray = 1
    unvignetted: bool | na.ScalarArray = True

    @property
    def shape(self) -> dict[str, int]:
        return na.shape_broadcasted(
            self.wavelength,
            self.position,
            self.direction,
            self.intensity,
            self.attenuation,
            self.index_refraction,
            self.unvignetted,
        )

    @property
    def n(self):
        """Complex index ``n + i alpha lambda / 4 pi`` (``_ray_vectors.py:74-88``)."""
        return self.index_refraction + self.attenuation * self.wavelength / (4 * np.pi) * 1j

    def copy_shallow(self) -> "RayVectorArray":
        return dataclasses.replace(self)

    def replace(self, **kwargs) -> "RayVectorArray":
        return dataclasses.replace(self, **kwargs)

    def __getitem__(self, item: dict) -> "RayVectorArray":
        def index(a):
            if isinstance(a, na.Cartesian3dVectorArray):
                return a[item]
            if isinstance(a, na.ScalarArray):
                return a[{ax: i for ax, i in item.items() if ax in a.axes}]
            return a

        return RayVectorArray(
            **{f.name: index(getattr(self, f.name)) for f in dataclasses.fields(self)}
        )


@dataclasses.dataclass(eq=False)
class RayFunctionArray(na.FunctionArray):
    """Rays (`outputs`) as a function of the object grid (`inputs`), ``optika/rays/_ray_functions.py``."""

    def __getitem__(self, item: dict) -> "RayFunctionArray":
        return RayFunctionArray(inputs=self.inputs, outputs=self.outputs[item])
