"""
Diffraction-grating rulings.

Host-side descriptions with the field names of ``optika.rulings`` (reference:
``optika/rulings/_rulings.py:131-251`` and ``optika/rulings/_spacing.py``).
The grating equation is applied inside the fused CUDA kernel as an
"effective incident direction" that is then fed to Snell's law, exactly as the
reference does (``optika/rulings/_rulings.py:24-128``, ``optika/surfaces.py:150-154``).
"""

from __future__ import annotations
import dataclasses
from . import named as na
from .transformations import AbstractTransformation

__all__ = [
    "AbstractRulingSpacing",
    "ConstantRulingSpacing",
    "Polynomial1dRulingSpacing",
    "HolographicRulingSpacing",
    "AbstractRulings",
    "Rulings",
    "MeasuredRulings",
    "SinusoidalRulings",
    "SquareRulings",
    "SawtoothRulings",
    "TriangularRulings",
    "RectangularRulings",
]


@dataclasses.dataclass(eq=False)
class AbstractRulingSpacing:
    """Interface: local ruling vector at a point (``optika/rulings/_spacing.py:14-41``)."""

    def __call__(self, position: na.Cartesian3dVectorArray, normal: na.Cartesian3dVectorArray):
        from . import _engine

        return _engine.ruling_vector(self, position, normal)


@dataclasses.dataclass(eq=False)
class ConstantRulingSpacing(AbstractRulingSpacing):
    """Constant spacing (``optika/rulings/_spacing.py:45-74``)."""

    constant: float | na.ScalarArray = 0
    normal: na.Cartesian3dVectorArray = dataclasses.field(
        default_factory=lambda: na.Cartesian3dVectorArray(1, 0, 0)
    )

    @property
    def transformation(self) -> None:
        return None

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(na.shape(self.constant), na.shape(self.normal))


@dataclasses.dataclass(eq=False)
class Polynomial1dRulingSpacing(AbstractRulingSpacing):
    """
    Variable line spacing ``d(x) = sum_k c_k x^k`` with ``x = position . normal``
    (``optika/rulings/_spacing.py:78-128``).  `coefficients` maps the integer
    power ``k`` to ``c_k`` in mm^(1-k); `transformation` is applied (forwards) to
    the position before the polynomial is evaluated.
    """

    coefficients: dict[int, float | na.ScalarArray] = None
    normal: na.Cartesian3dVectorArray = dataclasses.field(
        default_factory=lambda: na.Cartesian3dVectorArray(1, 0, 0)
    )
    transformation: None | AbstractTransformation = None

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(
            *[na.shape(c) for c in self.coefficients.values()],
            na.shape(self.normal),
            na.shape(self.transformation),
        )


@dataclasses.dataclass(eq=False)
class HolographicRulingSpacing(AbstractRulingSpacing):
    """
    Rulings recorded by the interference of two beams originating at `x1`, `x2`
    (``optika/rulings/_spacing.py:132-328``).  As in the reference, the declared
    `transformation` field is not applied by ``__call__`` (``_spacing.py:295-328``).
    """

    x1: na.Cartesian3dVectorArray = None
    x2: na.Cartesian3dVectorArray = None
    wavelength: float | na.ScalarArray = 0
    is_diverging_1: bool | na.ScalarArray = True
    is_diverging_2: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(
            na.shape(self.x1),
            na.shape(self.x2),
            na.shape(self.wavelength),
            na.shape(self.is_diverging_1),
            na.shape(self.is_diverging_2),
        )


@dataclasses.dataclass(eq=False)
class AbstractRulings:
    """Interface of a ruled surface (``optika/rulings/_rulings.py:131-221``)."""

    @property
    def spacing_(self) -> AbstractRulingSpacing:
        """`spacing` normalised to an :class:`AbstractRulingSpacing` (``_rulings.py:156-168``)."""
        spacing = self.spacing
        if not isinstance(spacing, AbstractRulingSpacing):
            spacing = ConstantRulingSpacing(
                constant=spacing,
                normal=na.Cartesian3dVectorArray(1, 0, 0),
            )
        return spacing

    def incident_effective(self, rays, normal: na.Cartesian3dVectorArray):
        """Effective incident direction (``optika/rulings/_rulings.py:170-204``)."""
        from . import _engine

        return _engine.rulings_incident_effective(self, rays, normal)

    def efficiency(self, rays, normal: na.Cartesian3dVectorArray):
        """Fraction of the light diffracted into the order (``optika/rulings/_rulings.py:205-221``), on the device."""
        from . import _engine

        return _engine.surface_efficiency(rays, normal, rulings=self)


@dataclasses.dataclass(eq=False)
class Rulings(AbstractRulings):
    """Ideal rulings with unit efficiency in every order (``_rulings.py:224-251``)."""

    spacing: float | na.ScalarArray | AbstractRulingSpacing = None
    diffraction_order: int | na.ScalarArray = 1

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(
            na.shape(self.spacing),
            na.shape(self.diffraction_order),
        )

    def efficiency(self, rays, normal) -> float:
        return 1


@dataclasses.dataclass(eq=False)
class MeasuredRulings(AbstractRulings):
    """
    Rulings whose efficiency was measured as a function of wavelength
    (``optika/rulings/_rulings.py:254-313``): ``numpy.interp`` of the ray wavelength in
    ``efficiency_measured`` (a :class:`~optika_b200.named.FunctionArray` whose inputs carry
    a one-dimensional ``wavelength`` and a single ``direction``).
    """

    spacing: float | na.ScalarArray | AbstractRulingSpacing = None
    diffraction_order: int | na.ScalarArray = 1
    efficiency_measured: na.FunctionArray = None

    @property
    def shape(self) -> dict[str, int]:
        shape_ = dict(na.shape(self.efficiency_measured.outputs))
        for ax in na.shape(self.efficiency_measured.inputs.wavelength):
            shape_.pop(ax, None)
        return na.broadcast_shapes(na.shape(self.spacing), na.shape(self.diffraction_order), shape_)


@dataclasses.dataclass(eq=False)
class _ProfileRulings(AbstractRulings):
    spacing: float | na.ScalarArray | AbstractRulingSpacing = None
    depth: float | na.ScalarArray = 0
    diffraction_order: int | na.ScalarArray = 1

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(
            na.shape(self.spacing), na.shape(self.depth), na.shape(self.diffraction_order)
        )


@dataclasses.dataclass(eq=False)
class SinusoidalRulings(_ProfileRulings):
    """Sinusoidal groove profile, efficiency ``J_m(2 gamma)`` (``optika/rulings/_rulings.py:316-457``)."""


@dataclasses.dataclass(eq=False)
class SquareRulings(_ProfileRulings):
    """Square-wave groove profile (``optika/rulings/_rulings.py:460-614``)."""


@dataclasses.dataclass(eq=False)
class SawtoothRulings(_ProfileRulings):
    """Sawtooth (blazed) groove profile (``optika/rulings/_rulings.py:617-758``)."""


@dataclasses.dataclass(eq=False)
class TriangularRulings(_ProfileRulings):
    """Triangular groove profile (``optika/rulings/_rulings.py:761-911``)."""


@dataclasses.dataclass(eq=False)
class RectangularRulings(_ProfileRulings):
    """Rectangular groove profile with a duty cycle (``optika/rulings/_rulings.py:914-1073``)."""

    ratio_duty: float | na.ScalarArray = 0.5

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(super().shape, na.shape(self.ratio_duty))
