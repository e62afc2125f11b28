"""
Optical materials: what happens to a ray when it meets a surface.

Mirrors ``optika.materials`` (``optika/materials/_materials.py``): `Vacuum`
(``:82-116``), `Mirror` (``:120-175``), `Glass` with the three-term Sellmeier
equation (``:309-455``).  The classes hold parameters; index of refraction,
attenuation, efficiency and the mirror flag are evaluated per ray inside the
fused CUDA kernel.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from .. import named as na
from .. import units as u

__all__ = [
    "AbstractMaterial",
    "Vacuum",
    "AbstractMirror",
    "Mirror",
    "MeasuredMirror",
    "Glass",
    "AbstractMultilayerMaterial",
    "MultilayerFilm",
    "MultilayerMirror",
]


@dataclasses.dataclass(eq=False)
class AbstractMaterial:
    """Interface of an optical material (``_materials.py:24-79``)."""

    @property
    def transformation(self) -> None:
        return None

    @property
    def shape(self) -> dict[str, int]:
        return {}

    @property
    def is_mirror(self) -> bool:
        return False


@dataclasses.dataclass(eq=False)
class Vacuum(AbstractMaterial):
    """Empty space: n = 1, no attenuation, unit efficiency (``_materials.py:82-116``)."""


@dataclasses.dataclass(eq=False)
class AbstractMirror(AbstractMaterial):
    """Reflects; index and attenuation pass through (``_materials.py:120-157``)."""

    @property
    def is_mirror(self) -> bool:
        return True


@dataclasses.dataclass(eq=False)
class Mirror(AbstractMirror):
    """An ideal mirror with unit efficiency (``_materials.py:160-175``)."""

    substrate: object = None


@dataclasses.dataclass(eq=False)
class MeasuredMirror(AbstractMirror):
    """
    A mirror whose reflectivity was measured as a function of wavelength
    (``optika/materials/_materials.py:177-305``): ``efficiency_measured`` is a
    :class:`~optika_b200.named.FunctionArray` with inputs
    :class:`~optika_b200.vectors.SpectralDirectionalVectorArray` (one-dimensional
    ``wavelength``, a single ``direction``) and the reflectivity as outputs; the efficiency
    of a ray is ``numpy.interp`` of its wavelength in that table (``:301-305``), evaluated
    inside the trace kernel.
    """

    efficiency_measured: na.FunctionArray = None
    substrate: object = None
    serial_number: object = None

    @property
    def shape(self) -> dict[str, int]:
        shape_ = dict(na.shape(self.efficiency_measured.outputs))
        for ax in na.shape(self.efficiency_measured.inputs.wavelength):
            shape_.pop(ax, None)
        return shape_

    def efficiency(self, rays, normal):
        from .. import _engine

        return _engine.surface_efficiency(rays, normal, material=self)


@dataclasses.dataclass(eq=False)
class Glass(AbstractMaterial):
    """
    Transparent glass following the three-term Sellmeier equation
    (``_materials.py:309-455``).  ``c1..c3`` are in mm^2 (use ``u.um ** 2``).
    As in the reference, the equation is evaluated on ``rays.wavelength``,
    which is the in-medium wavelength after a previous refraction.
    """

    b1: float | na.ScalarArray = 0
    b2: float | na.ScalarArray = 0
    b3: float | na.ScalarArray = 0
    c1: float | na.ScalarArray = 0
    c2: float | na.ScalarArray = 0
    c3: float | na.ScalarArray = 0

    @classmethod
    def n_bk7(cls) -> "Glass":
        """SCHOTT N-BK7 (``_materials.py:382-396``)."""
        return cls(
            b1=1.03961212,
            b2=0.231792344,
            b3=1.01046945,
            c1=0.00600069867 * u.um**2,
            c2=0.0200179144 * u.um**2,
            c3=103.560653 * u.um**2,
        )

    @classmethod
    def f2(cls) -> "Glass":
        """SCHOTT F2 (``_materials.py:398-412``)."""
        return cls(
            b1=1.34533359,
            b2=0.209073176,
            b3=0.937357162,
            c1=0.00997743871 * u.um**2,
            c2=0.0470450767 * u.um**2,
            c3=111.886764 * u.um**2,
        )

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(
            *[na.shape(p) for p in (self.b1, self.b2, self.b3, self.c1, self.c2, self.c3)]
        )


@dataclasses.dataclass(eq=False)
class AbstractMultilayerMaterial(AbstractMaterial):
    """
    A multilayer coating as a surface material (``optika/materials/_multilayers.py:772-826``):
    index of refraction and attenuation pass through; the efficiency of every ray is
    ``multilayer_efficiency`` at its wavelength and cosine of incidence, evaluated on the
    device (``_engine.trace`` chains the launches around the coated surface).
    """

    @property
    def _substrate(self):
        return None

    @property
    def _polarization_rows(self) -> tuple[int, int]:
        raise NotImplementedError

    @property
    def shape(self) -> dict[str, int]:
        from ._multilayers import flatten_layers

        flat, _ = flatten_layers(self.layers)
        parts = [na.shape(layer) for layer in flat]
        if self._substrate is not None:
            parts.append(na.shape(self._substrate))
        return na.broadcast_shapes(*parts)

    def efficiency_device(self, wavelength, cos_incidence, index_refraction, attenuation, config_shape, cindex, device):
        """Device tensor ``[2, N]``: the s and p efficiencies of N rays (R for mirrors, T for films)."""
        from ._multilayers import multilayer_efficiency_rays

        out = multilayer_efficiency_rays(
            wavelength, cos_incidence, index_refraction, attenuation, self.layers, self._substrate,
            config_shape, cindex, device,
        )
        a, b = self._polarization_rows
        return out[a : b + 1]

    def efficiency(self, rays, normal):
        """``efficiency(rays, normal)`` of the reference (``:839-866, 908-935``) for host rays."""
        from ._multilayers import multilayer_efficiency

        wavelength = u.length(rays.wavelength)
        k = rays.attenuation * wavelength / (4 * np.pi)
        n = rays.index_refraction + k * 1j
        reflectivity, transmissivity = multilayer_efficiency(
            wavelength=wavelength, direction=-(rays.direction @ normal), n=n, layers=self.layers,
            substrate=self._substrate,
        )
        return (reflectivity if self.is_mirror else transmissivity).average


@dataclasses.dataclass(eq=False)
class MultilayerFilm(AbstractMultilayerMaterial):
    """A free-standing thin-film stack: efficiency = transmissivity (``_multilayers.py:829-895``)."""

    layers: object = None

    @property
    def _polarization_rows(self) -> tuple[int, int]:
        return 2, 3  # T_s, T_p


@dataclasses.dataclass(eq=False)
class MultilayerMirror(AbstractMultilayerMaterial, AbstractMirror):
    """A multilayer-coated mirror: efficiency = reflectivity (``_multilayers.py:898-987``)."""

    layers: object = None
    substrate: object = None

    @property
    def _substrate(self):
        return self.substrate

    @property
    def _polarization_rows(self) -> tuple[int, int]:
        return 0, 1  # R_s, R_p
