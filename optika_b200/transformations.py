"""
Rigid coordinate transformations with the spellings of ``named_arrays.transformations``.

The reference positions every surface, sag, aperture and ruling pattern with a
``na.transformations.AbstractTransformation`` and applies it to rays as
"position -> R p + t, direction -> R d", and its inverse as
"position -> R^T (p - t), direction -> R^T d"
(call sites ``optika/surfaces.py:141-142, 195-196``; dispatch
``optika/rays/_ray_vectors.py:105-166``).  ``named_arrays`` is a third-party
dependency that is absent here; its conventions are pinned by
``optika/_util_test.py:20-32`` (right-handed rotation matrices) and by the
geometry of the Newtonian example (``optika/systems/_sequential.py:1882-1912``:
a ``TransformationList`` applies its first element first).

Every transformation reduces to an :class:`Affine` (3x3 matrix `R`, vector `t`)
whose entries may carry named configuration axes; the lowering pass
(:mod:`optika_b200._lowering`) evaluates it per configuration and packs it into
the device surface table.
"""

from __future__ import annotations
from typing import Sequence
import dataclasses
import numpy as np
from . import named as na
from . import units as u

__all__ = [
    "Affine",
    "AbstractTransformation",
    "IdentityTransformation",
    "Cartesian3dTranslation",
    "Translation",
    "Cartesian3dRotationX",
    "Cartesian3dRotationY",
    "Cartesian3dRotationZ",
    "TransformationList",
]


@dataclasses.dataclass(eq=False)
class Affine:
    """``x -> R x + t``.  `matrix` is a 3x3 nested tuple, `vector` a 3-tuple."""

    matrix: tuple
    vector: tuple

    @classmethod
    def identity(cls) -> "Affine":
        return cls(
            matrix=((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)),
            vector=(0.0, 0.0, 0.0),
        )

    def __matmul__(self, other: "Affine") -> "Affine":
        """Composition: ``(self @ other)(x) = self(other(x))``."""
        a, b = self.matrix, other.matrix
        m = tuple(
            tuple(sum(a[i][k] * b[k][j] for k in range(3)) for j in range(3))
            for i in range(3)
        )
        v = tuple(
            sum(a[i][k] * other.vector[k] for k in range(3)) + self.vector[i]
            for i in range(3)
        )
        return Affine(m, v)

    @property
    def inverse(self) -> "Affine":
        """Inverse of a rigid motion: ``x -> R^T (x - t)``."""
        a = self.matrix
        m = tuple(tuple(a[j][i] for j in range(3)) for i in range(3))
        v = tuple(-sum(m[i][k] * self.vector[k] for k in range(3)) for i in range(3))
        return Affine(m, v)

    @property
    def shape(self) -> dict[str, int]:
        entries = [e for row in self.matrix for e in row] + list(self.vector)
        return na.shape_broadcasted(*entries)

    def numpy(self, shape_: dict[str, int]) -> tuple[np.ndarray, np.ndarray]:
        """Dense ``R[..., 3, 3]`` and ``t[..., 3]`` broadcast over `shape_`."""
        dims = tuple(shape_.values())
        r = np.empty(dims + (3, 3))
        t = np.empty(dims + (3,))
        for i in range(3):
            for j in range(3):
                r[..., i, j] = np.broadcast_to(na.aligned(self.matrix[i][j], shape_), dims)
            t[..., i] = np.broadcast_to(na.aligned(self.vector[i], shape_), dims)
        return r, t

    def apply_position(self, p: na.Cartesian3dVectorArray) -> na.Cartesian3dVectorArray:
        m, v = self.matrix, self.vector
        c = (p.x, p.y, p.z)
        return na.Cartesian3dVectorArray(
            *[sum(m[i][k] * c[k] for k in range(3)) + v[i] for i in range(3)]
        )

    def apply_direction(self, d: na.Cartesian3dVectorArray) -> na.Cartesian3dVectorArray:
        m = self.matrix
        c = (d.x, d.y, d.z)
        return na.Cartesian3dVectorArray(
            *[sum(m[i][k] * c[k] for k in range(3)) for i in range(3)]
        )


class AbstractTransformation:
    """Interface shared by all rigid transformations."""

    @property
    def affine(self) -> Affine:
        raise NotImplementedError

    @property
    def shape(self) -> dict[str, int]:
        return self.affine.shape

    @property
    def inverse(self) -> "AbstractTransformation":
        return _Composed(self.affine.inverse)

    def __matmul__(self, other: "AbstractTransformation") -> "AbstractTransformation":
        return _Composed(self.affine @ other.affine)

    def __call__(self, a):
        """Apply to a 3-D position vector, or to rays (position affine, direction linear)."""
        aff = self.affine
        if isinstance(a, na.Cartesian3dVectorArray):
            return aff.apply_position(a)
        if hasattr(a, "position") and hasattr(a, "direction"):
            return dataclasses.replace(
                a,
                position=aff.apply_position(a.position),
                direction=aff.apply_direction(a.direction),
            )
        raise TypeError(f"cannot transform object of type {type(a)}")


@dataclasses.dataclass(eq=False)
class _Composed(AbstractTransformation):
    _affine: Affine

    @property
    def affine(self) -> Affine:
        return self._affine


@dataclasses.dataclass(eq=False)
class IdentityTransformation(AbstractTransformation):
    @property
    def affine(self) -> Affine:
        return Affine.identity()


@dataclasses.dataclass(eq=False)
class Cartesian3dTranslation(AbstractTransformation):
    """Translation by ``(x, y, z)`` (engine length units, mm)."""

    x: float | na.ScalarArray = 0
    y: float | na.ScalarArray = 0
    z: float | na.ScalarArray = 0

    @property
    def affine(self) -> Affine:
        return Affine(
            matrix=Affine.identity().matrix,
            vector=(u.length(self.x), u.length(self.y), u.length(self.z)),
        )


def Translation(vector: na.Cartesian3dVectorArray) -> Cartesian3dTranslation:
    """``na.transformations.Translation`` for a 3-D displacement vector."""
    return Cartesian3dTranslation(vector.x, vector.y, vector.z)


@dataclasses.dataclass(eq=False)
class Cartesian3dRotationX(AbstractTransformation):
    """Right-handed rotation about the x axis by `angle` (radians)."""

    angle: float | na.ScalarArray = 0

    @property
    def affine(self) -> Affine:
        a = u.angle(self.angle)
        c, s = np.cos(a), np.sin(a)
        return Affine(((1.0, 0.0, 0.0), (0.0, c, -s), (0.0, s, c)), (0.0, 0.0, 0.0))


@dataclasses.dataclass(eq=False)
class Cartesian3dRotationY(AbstractTransformation):
    """Right-handed rotation about the y axis by `angle` (radians)."""

    angle: float | na.ScalarArray = 0

    @property
    def affine(self) -> Affine:
        a = u.angle(self.angle)
        c, s = np.cos(a), np.sin(a)
        return Affine(((c, 0.0, s), (0.0, 1.0, 0.0), (-s, 0.0, c)), (0.0, 0.0, 0.0))


@dataclasses.dataclass(eq=False)
class Cartesian3dRotationZ(AbstractTransformation):
    """Right-handed rotation about the z axis by `angle` (radians)."""

    angle: float | na.ScalarArray = 0

    @property
    def affine(self) -> Affine:
        a = u.angle(self.angle)
        c, s = np.cos(a), np.sin(a)
        return Affine(((c, -s, 0.0), (s, c, 0.0), (0.0, 0.0, 1.0)), (0.0, 0.0, 0.0))


@dataclasses.dataclass(eq=False)
class TransformationList(AbstractTransformation):
    """A sequence of transformations; the first element is applied first."""

    transformations: Sequence[AbstractTransformation] = ()

    @property
    def affine(self) -> Affine:
        result = Affine.identity()
        for t in self.transformations:
            if t is not None:
                result = t.affine @ result
        return result

    def __iter__(self):
        return iter(self.transformations)
