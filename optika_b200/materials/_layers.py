"""
Layers of a multilayer stack.

Mirrors ``optika.materials.Layer``, ``LayerSequence`` and
``PeriodicLayerSequence`` (``optika/materials/_layers.py:131-277``, ``:377-499``,
``:503-645``).  The classes hold parameters; the transfer-matrix chain is
evaluated by ``optika_b200/csrc/multilayer.cu``.
"""

from __future__ import annotations
from typing import Sequence
import dataclasses
from .. import named as na
from .. import chemicals
from .profiles import AbstractInterfaceProfile

__all__ = ["AbstractLayer", "Layer", "LayerSequence", "PeriodicLayerSequence"]


class AbstractLayer:
    pass


@dataclasses.dataclass(eq=False)
class Layer(AbstractLayer):
    """A homogeneous layer (``_layers.py:131-277``).  `chemical` None means vacuum."""

    chemical: None | str | chemicals.AbstractChemical = None
    thickness: None | float | na.ScalarArray = None
    interface: None | AbstractInterfaceProfile = None

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(na.shape(self.thickness), na.shape(self.interface))

    @property
    def _chemical(self) -> chemicals.AbstractChemical:
        result = self.chemical
        if not isinstance(result, chemicals.AbstractChemical):
            result = chemicals.Chemical(result)
        return result

    def n(self, wavelength):
        """Complex index of refraction of the layer medium (``_layers.py:218-227``)."""
        if self.chemical is None:
            return 1
        return self._chemical.n(wavelength)

    @property
    def layer_sequence(self) -> "LayerSequence":
        return LayerSequence([self])


@dataclasses.dataclass(eq=False)
class LayerSequence(AbstractLayer):
    """An explicit sequence of layers (``_layers.py:377-499``)."""

    layers: Sequence[AbstractLayer] = ()

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(*[na.shape(layer) for layer in self.layers])


@dataclasses.dataclass(eq=False)
class PeriodicLayerSequence(AbstractLayer):
    """`layers` repeated `num_periods` times (``_layers.py:503-645``)."""

    layers: Sequence[AbstractLayer] = ()
    num_periods: int = 1

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(*[na.shape(layer) for layer in self.layers])

    @property
    def layer_sequence(self) -> LayerSequence:
        return LayerSequence(list(self.layers) * self.num_periods)
