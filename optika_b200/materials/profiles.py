"""
Interface profiles between layers of a multilayer stack.

Mirrors ``optika.materials.profiles`` (``optika/materials/profiles.py``).  Each
profile multiplies the Fresnel reflection coefficient of its interface by the
Fourier transform of the derivative of the profile, evaluated at
``s = Re(4 pi n cos(theta) / lambda)`` (``profiles.py:103-126``):

* erf (``:221-222``):          ``exp(-(s w)^2 / 2)``
* exponential (``:320-325``):  ``1 / (1 + (s w)^2 / 2)``
* linear (``:424-431``):       ``sin(sqrt(3) w s) / (sqrt(3) w s)``
* sinusoidal (``:532-541``):   ``pi/4 (sin(x - pi/2)/(x - pi/2) + sin(x + pi/2)/(x + pi/2))``,
  ``x = a w s``, ``a = pi / (pi^2 - 8)``

The formulas are evaluated in ``optika_b200/csrc/multilayer.cu``.
"""

from __future__ import annotations
import dataclasses
from .. import named as na

__all__ = [
    "AbstractInterfaceProfile",
    "ErfInterfaceProfile",
    "ExponentialInterfaceProfile",
    "LinearInterfaceProfile",
    "SinusoidalInterfaceProfile",
]


@dataclasses.dataclass(eq=False)
class AbstractInterfaceProfile:
    width: float | na.ScalarArray = 0
    """Characteristic width of the interface (engine length units, mm)."""

    kind = 0

    @property
    def shape(self) -> dict[str, int]:
        return na.shape(self.width)


@dataclasses.dataclass(eq=False)
class ErfInterfaceProfile(AbstractInterfaceProfile):
    kind = 1


@dataclasses.dataclass(eq=False)
class ExponentialInterfaceProfile(AbstractInterfaceProfile):
    kind = 2


@dataclasses.dataclass(eq=False)
class LinearInterfaceProfile(AbstractInterfaceProfile):
    kind = 3


@dataclasses.dataclass(eq=False)
class SinusoidalInterfaceProfile(AbstractInterfaceProfile):
    kind = 4
