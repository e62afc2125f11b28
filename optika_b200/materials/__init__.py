"""Optical materials (mirrors ``optika.materials``)."""

from ._materials import (
    AbstractMaterial, Vacuum, AbstractMirror, Mirror, MeasuredMirror, Glass,
    AbstractMultilayerMaterial, MultilayerFilm, MultilayerMirror,
)
from . import profiles
from ._layers import AbstractLayer, Layer, LayerSequence, PeriodicLayerSequence
from ._snells_law import snells_law, snells_law_scalar
from ._multilayers import multilayer_efficiency

__all__ = [
    "AbstractMaterial",
    "Vacuum",
    "AbstractMirror",
    "Mirror",
    "MeasuredMirror",
    "AbstractMultilayerMaterial",
    "MultilayerFilm",
    "MultilayerMirror",
    "Glass",
    "profiles",
    "AbstractLayer",
    "Layer",
    "LayerSequence",
    "PeriodicLayerSequence",
    "snells_law",
    "snells_law_scalar",
    "multilayer_efficiency",
]
