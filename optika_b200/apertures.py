"""
Apertures and obscurations.

Host-side descriptions with the field names of ``optika.apertures``
(reference: ``optika/apertures/_apertures.py``).  The containment tests run in
the fused CUDA kernel; these classes only hold parameters and know how to
produce their polygon vertices and bounds.

Unit handling: the reference decides between clipping on ray *position* and
ray *direction* by the unit of ``bound_lower`` (``_apertures.py:82-102``:
lengths -> position, dimensionless -> direction).  Engine-side values are plain
floats, so an aperture that is specified in direction cosines says so with
``angular=True``.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na
from . import units as u
from .transformations import AbstractTransformation

__all__ = [
    "AbstractAperture",
    "CircularAperture",
    "CircularSectorAperture",
    "EllipticalAperture",
    "AbstractPolygonalAperture",
    "PolygonalAperture",
    "RectangularAperture",
    "RegularPolygonalAperture",
    "OctagonalAperture",
    "IsoscelesTrapezoidalAperture",
]


@dataclasses.dataclass(eq=False)
class AbstractAperture:
    """Interface of an aperture (``optika/apertures/_apertures.py:28-132``)."""

    @property
    def _parameters(self) -> tuple:
        return ()

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(
            *[na.shape(p) for p in self._parameters],
            na.shape(self.active),
            na.shape(self.inverted),
            na.shape(self.transformation),
        )

    def __call__(self, position: na.Cartesian3dVectorArray):
        """Boolean mask, True where `position` is inside (``_apertures.py:69-80``)."""
        from . import _engine

        return _engine.aperture_mask(self, position)

    def clip_rays(self, rays):
        """AND the containment test into ``rays.unvignetted`` (``_apertures.py:82-102``)."""
        from . import _engine

        return _engine.aperture_clip(self, rays)


@dataclasses.dataclass(eq=False)
class CircularAperture(AbstractAperture):
    """``sqrt(x^2 + y^2) <= radius`` (``_apertures.py:233-314``)."""

    radius: float | na.ScalarArray = 0
    active: bool | na.ScalarArray = True
    inverted: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None
    angular: bool = dataclasses.field(default=False, kw_only=True)

    @property
    def _parameters(self):
        return (self.radius,)

    @property
    def bound_lower(self) -> na.Cartesian3dVectorArray:
        r = u.length(self.radius)
        b = na.Cartesian3dVectorArray(-r, -r, 0 * r)
        return _bound(self, b, lower=True)

    @property
    def bound_upper(self) -> na.Cartesian3dVectorArray:
        r = u.length(self.radius)
        b = na.Cartesian3dVectorArray(r, r, 0 * r)
        return _bound(self, b, lower=False)

    def wire(self, num: int = 101) -> na.Cartesian3dVectorArray:
        """Points on the edge, on a new ``"wire"`` axis (``_apertures.py:337-356``)."""
        az = na.linspace(0, 360, axis="wire", num=num) * u.deg
        r = u.length(self.radius)
        result = na.Cartesian3dVectorArray(x=r * np.cos(az), y=r * np.sin(az), z=0 * az)
        if self.transformation is not None:
            result = self.transformation(result)
        return result


@dataclasses.dataclass(eq=False)
class CircularSectorAperture(AbstractAperture):
    """Circle restricted to an angular range (``_apertures.py:367-480``)."""

    radius: float | na.ScalarArray = 0
    angle_start: float | na.ScalarArray = 0
    angle_stop: float | na.ScalarArray = 180 * u.deg
    active: bool | na.ScalarArray = True
    inverted: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None
    angular: bool = dataclasses.field(default=False, kw_only=True)

    @property
    def _parameters(self):
        return (self.radius, self.angle_start, self.angle_stop)


@dataclasses.dataclass(eq=False)
class EllipticalAperture(AbstractAperture):
    """``(x/a)^2 + (y/b)^2 <= 1`` (``_apertures.py:571-663``)."""

    radius: na.Cartesian2dVectorArray = dataclasses.field(
        default_factory=lambda: na.Cartesian2dVectorArray(0, 0)
    )
    active: bool | na.ScalarArray = True
    inverted: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None
    angular: bool = dataclasses.field(default=False, kw_only=True)

    @property
    def _parameters(self):
        return (self.radius.x, self.radius.y)


def _bound(aperture, b: na.Cartesian3dVectorArray, lower: bool):
    t = aperture.transformation
    if t is None:
        return b
    return t(b)


@dataclasses.dataclass(eq=False)
class AbstractPolygonalAperture(AbstractAperture):
    """Polygon described by `vertices` on a ``"vertex"`` axis (``_apertures.py:716-856``)."""

    @property
    def vertices(self) -> na.Cartesian3dVectorArray:
        raise NotImplementedError

    @property
    def shape(self) -> dict[str, int]:
        s = na.broadcast_shapes(
            *[na.shape(p) for p in self._parameters],
            na.shape(self.active),
            na.shape(self.inverted),
            na.shape(self.transformation),
        )
        s.pop("vertex", None)
        return s

    @property
    def bound_lower(self) -> na.Cartesian3dVectorArray:
        v = self.vertices
        if self.transformation is not None:
            v = self.transformation(v)
        return v.min(axis="vertex")

    @property
    def bound_upper(self) -> na.Cartesian3dVectorArray:
        v = self.vertices
        if self.transformation is not None:
            v = self.transformation(v)
        return v.max(axis="vertex")

    def wire(self, num: int = 101) -> na.Cartesian3dVectorArray:
        """`num` points along the polygon's sides on a new ``"wire"`` axis (``_apertures.py:794-836``)."""
        v = self.vertices
        shape_ = na.shape(v)
        n = shape_["vertex"]
        vx = na.broadcast_to(na.as_named_array(v.x), shape_)
        vy = na.broadcast_to(na.as_named_array(v.y), shape_)
        vz = na.broadcast_to(na.as_named_array(v.z), shape_)
        per_side = num / n
        pieces = []
        cumulative = 0
        for k in range(n):
            num_k = int((k + 1) * per_side - cumulative)
            cumulative += num_k
            endpoint = cumulative == num
            t = na.linspace(0, 1, axis="wire", num=num_k, endpoint=endpoint)
            left = [c[dict(vertex=k)] for c in (vx, vy, vz)]
            right = [c[dict(vertex=(k + 1) % n)] for c in (vx, vy, vz)]
            pieces.append([a + (b - a) * t for a, b in zip(left, right)])
        comps = []
        for c in range(3):
            arrays = [p[c] for p in pieces]
            shp = na.shape_broadcasted(*[a[dict(wire=0)] for a in arrays])
            nd = np.concatenate(
                [np.broadcast_to(na.aligned(a, dict(wire=a.shape["wire"], **shp)), (a.shape["wire"],) + tuple(shp.values())) for a in arrays],
                axis=0,
            )
            comps.append(na.ScalarArray(nd, ("wire",) + tuple(shp)))
        result = na.Cartesian3dVectorArray(*comps)
        if self.transformation is not None:
            result = self.transformation(result)
        return result


@dataclasses.dataclass(eq=False)
class PolygonalAperture(AbstractPolygonalAperture):
    """Arbitrary polygon (``_apertures.py:839-856``)."""

    vertices: na.Cartesian3dVectorArray = None
    active: bool | na.ScalarArray = True
    inverted: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None
    angular: bool = dataclasses.field(default=False, kw_only=True)

    @property
    def _parameters(self):
        return (self.vertices.x, self.vertices.y)


@dataclasses.dataclass(eq=False)
class RectangularAperture(AbstractPolygonalAperture):
    """``-h <= (x, y) <= h``, inclusive on both sides (``_apertures.py:859-986``)."""

    half_width: float | na.ScalarArray | na.Cartesian2dVectorArray = 0
    active: bool | na.ScalarArray = True
    inverted: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None
    angular: bool = dataclasses.field(default=False, kw_only=True)

    @property
    def half_width_xy(self) -> na.Cartesian2dVectorArray:
        h = self.half_width
        if isinstance(h, na.Cartesian2dVectorArray):
            return na.Cartesian2dVectorArray(u.length(h.x), u.length(h.y))
        h = u.length(h)
        return na.Cartesian2dVectorArray(h, h)

    @property
    def _parameters(self):
        h = self.half_width_xy
        return (h.x, h.y)

    @property
    def vertices(self) -> na.Cartesian3dVectorArray:
        # optika/apertures/_apertures.py:970-986: sqrt(2) * (cos, sin)(45deg + k 90deg) * half_width
        h = self.half_width_xy
        # degrees first, one conversion to radians (as astropy does inside np.cos)
        az = (na.linspace(0, 360, axis="vertex", num=4, endpoint=False) + 45) * u.deg
        r = np.sqrt(2)
        return na.Cartesian3dVectorArray(
            x=r * np.cos(az) * h.x,
            y=r * np.sin(az) * h.y,
            z=0 * az,
        )


@dataclasses.dataclass(eq=False)
class RegularPolygonalAperture(AbstractPolygonalAperture):
    """Regular polygon with vertices at `radius` (``_apertures.py:989-1049``)."""

    radius: float | na.ScalarArray = 0
    num_vertices: int = 0
    active: bool | na.ScalarArray = True
    inverted: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None
    angular: bool = dataclasses.field(default=False, kw_only=True)

    @property
    def _parameters(self):
        return (self.radius,)

    @property
    def vertices(self) -> na.Cartesian3dVectorArray:
        # optika/apertures/_apertures.py:1009-1027
        radius = u.length(self.radius)
        angle = na.linspace(0, 360, axis="vertex", num=self.num_vertices, endpoint=False) * u.deg
        return na.Cartesian3dVectorArray(
            x=radius * np.cos(angle),
            y=radius * np.sin(angle),
            z=0 * angle,
        )


@dataclasses.dataclass(eq=False)
class OctagonalAperture(RegularPolygonalAperture):
    """Regular octagon (``_apertures.py:1052-1079``)."""

    num_vertices: int = 8


@dataclasses.dataclass(eq=False)
class IsoscelesTrapezoidalAperture(AbstractPolygonalAperture):
    """Isosceles trapezoid between `x_left` and `x_right` (``_apertures.py:1082-1193``)."""

    x_left: float | na.ScalarArray = 0
    x_right: float | na.ScalarArray = 0
    angle: float | na.ScalarArray = 0
    active: bool | na.ScalarArray = True
    inverted: bool | na.ScalarArray = False
    transformation: None | AbstractTransformation = None
    angular: bool = dataclasses.field(default=False, kw_only=True)

    @property
    def _parameters(self):
        return (self.x_left, self.x_right, self.angle)

    @property
    def vertices(self) -> na.Cartesian3dVectorArray:
        # optika/apertures/_apertures.py:1109-1134: (left, right) upper, then mirrored lower reversed
        x_left = u.length(self.x_left)
        x_right = u.length(self.x_right)
        m = np.tan(u.angle(self.angle) / 2)
        xs = [x_left, x_right, x_right, x_left]
        ys = [m * x_left, m * x_right, -(m * x_right), -(m * x_left)]
        return na.Cartesian3dVectorArray(
            x=na.stack([na.as_named_array(v) for v in xs], axis="vertex"),
            y=na.stack([na.as_named_array(v) for v in ys], axis="vertex"),
            z=na.stack([na.as_named_array(0 * v) for v in xs], axis="vertex"),
        )
