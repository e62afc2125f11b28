"""
A small named-axis array layer for the host side of the engine.

The reference does all array arithmetic through the third-party package
``named_arrays`` (``import named_arrays as na`` in every module, e.g.
``optika/surfaces.py:11-15``), which broadcasts by *axis name* rather than by
position.  That package is not vendored with the reference and is not available
where this engine is built, so this module provides the subset the hot path
needs, with the same spellings:

``ScalarArray(ndarray, axes)``, ``Cartesian2dVectorArray``, ``Cartesian3dVectorArray``,
``linspace``, ``arange``, ``stack``, ``shape``, ``broadcast_shapes``,
``shape_broadcasted``, ``broadcast_to``, ``Cartesian2dVectorLinearSpace``.

Nothing here runs on the hot path: these containers only describe ray grids
(field x pupil x wavelength x configuration) until :mod:`optika_b200._flatten`
turns them into strided structure-of-arrays buffers for the device.
"""

from __future__ import annotations
from typing import Sequence
import dataclasses
import numpy as np

__all__ = [
    "ScalarArray",
    "Cartesian2dVectorArray",
    "Cartesian3dVectorArray",
    "Cartesian2dVectorLinearSpace",
    "FunctionArray",
    "as_named_array",
    "shape",
    "broadcast_shapes",
    "shape_broadcasted",
    "broadcast_to",
    "linspace",
    "arange",
    "stack",
    "aligned",
]


def broadcast_shapes(*shapes: dict[str, int]) -> dict[str, int]:
    """Broadcast dictionaries of ``{axis: size}`` by axis name."""
    result: dict[str, int] = {}
    for s in shapes:
        for ax, n in s.items():
            if ax in result:
                if result[ax] == 1:
                    result[ax] = n
                elif n != 1 and result[ax] != n:
                    raise ValueError(
                        f"axis {ax!r} has incompatible sizes {result[ax]} and {n}"
                    )
            else:
                result[ax] = n
    return result


def shape(a) -> dict[str, int]:
    """The named shape of `a` (empty for plain scalars)."""
    if isinstance(a, (ScalarArray, _VectorArray)):
        return a.shape
    if hasattr(a, "shape") and isinstance(getattr(a, "shape"), dict):
        return a.shape
    return {}


def shape_broadcasted(*arrays) -> dict[str, int]:
    return broadcast_shapes(*[shape(a) for a in arrays])


def as_named_array(a) -> "ScalarArray":
    if isinstance(a, ScalarArray):
        return a
    a = np.asarray(a)
    if a.ndim != 0:
        raise ValueError("plain numpy arrays with dimensions need explicit axes")
    return ScalarArray(a, ())


def aligned(a, shape_: dict[str, int]) -> np.ndarray:
    """
    View of `a` as a plain ndarray whose dimensions follow the order of `shape_`,
    with size-1 dimensions for axes `a` does not have (ready for numpy broadcasting).
    """
    a = as_named_array(a)
    nd = np.asarray(a.ndarray)
    missing = [ax for ax in a.axes if ax not in shape_]
    if missing:
        raise ValueError(f"axes {missing} are not in the target shape {shape_}")
    order = [a.axes.index(ax) for ax in shape_ if ax in a.axes]
    nd = nd.transpose(order) if order else nd
    index = tuple(slice(None) if ax in a.axes else None for ax in shape_)
    return nd[index] if index else nd


def broadcast_to(a, shape_: dict[str, int]):
    if isinstance(a, _VectorArray):
        return a.broadcast_to(shape_)
    nd = aligned(a, shape_)
    nd = np.broadcast_to(nd, tuple(shape_.values()))
    return ScalarArray(nd, tuple(shape_))


class ScalarArray(np.lib.mixins.NDArrayOperatorsMixin):
    """An n-dimensional array whose dimensions are identified by name."""

    __array_priority__ = 1000

    def __init__(self, ndarray=0, axes: None | str | Sequence[str] = None):
        if isinstance(ndarray, ScalarArray):
            if axes is None:
                axes = ndarray.axes
            ndarray = ndarray.ndarray
        nd = np.asarray(ndarray)
        if axes is None:
            axes = ()
        elif isinstance(axes, str):
            axes = (axes,)
        axes = tuple(axes)
        if nd.ndim != len(axes):
            raise ValueError(
                f"number of axes {axes} does not match array dimensions {nd.shape}"
            )
        self.ndarray = nd
        self.axes = axes

    # -- basic properties -------------------------------------------------
    @property
    def shape(self) -> dict[str, int]:
        return dict(zip(self.axes, self.ndarray.shape))

    @property
    def ndim(self) -> int:
        return self.ndarray.ndim

    @property
    def size(self) -> int:
        return self.ndarray.size

    @property
    def dtype(self):
        return self.ndarray.dtype

    @property
    def value(self) -> "ScalarArray":
        return self

    @property
    def explicit(self) -> "ScalarArray":
        return self

    def __repr__(self) -> str:
        return f"ScalarArray({self.ndarray!r}, axes={self.axes!r})"

    def __float__(self) -> float:
        return float(self.ndarray)

    def __bool__(self) -> bool:
        return bool(self.ndarray)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.ndarray, dtype=dtype)

    def copy(self) -> "ScalarArray":
        return ScalarArray(self.ndarray.copy(), self.axes)

    def astype(self, dtype) -> "ScalarArray":
        return ScalarArray(self.ndarray.astype(dtype), self.axes)

    def broadcast_to(self, shape_: dict[str, int]) -> "ScalarArray":
        return broadcast_to(self, shape_)

    def numpy(self, axes: Sequence[str]) -> np.ndarray:
        """Plain ndarray with dimensions in the order of `axes` (all must be given)."""
        shape_ = {ax: self.shape.get(ax, 1) for ax in axes}
        return aligned(self, shape_)

    # -- numpy protocol ---------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        if any(isinstance(x, _VectorArray) for x in inputs):
            return NotImplemented
        shape_ = shape_broadcasted(*inputs)
        args = [
            aligned(x, shape_) if isinstance(x, ScalarArray) else x for x in inputs
        ]
        result = ufunc(*args, **kwargs)
        axes = tuple(shape_)
        if isinstance(result, tuple):
            return tuple(ScalarArray(r, axes) for r in result)
        return ScalarArray(result, axes)

    def __array_function__(self, func, types, args, kwargs):
        if func is np.where:
            cond, a, b = args
            shape_ = shape_broadcasted(cond, a, b)
            r = np.where(
                aligned(cond, shape_), aligned(a, shape_), aligned(b, shape_)
            )
            return ScalarArray(r, tuple(shape_))
        if func in (np.real, np.imag, np.conj, np.abs, np.square, np.sqrt):
            return ScalarArray(func(args[0].ndarray), args[0].axes)
        if func in (np.all, np.any):
            return func(args[0].ndarray)
        if func is np.allclose:
            a, b = args[:2]
            shape_ = shape_broadcasted(a, b)
            return np.allclose(aligned(a, shape_), aligned(b, shape_), **kwargs)
        return NotImplemented

    # -- indexing ---------------------------------------------------------
    def __getitem__(self, item) -> "ScalarArray":
        if isinstance(item, ScalarArray):
            raise NotImplementedError("boolean-mask indexing is not supported")
        if not isinstance(item, dict):
            raise TypeError("index a ScalarArray with a dict of {axis: index}")
        index = []
        axes = []
        for ax in self.axes:
            if ax in item:
                i = item[ax]
                if isinstance(i, ScalarArray):
                    raise NotImplementedError("advanced indexing is not supported")
                index.append(i)
                if isinstance(i, slice):
                    axes.append(ax)
            else:
                index.append(slice(None))
                axes.append(ax)
        return ScalarArray(self.ndarray[tuple(index)], tuple(axes))

    # -- reductions -------------------------------------------------------
    def _reduce(self, func, axis):
        if axis is None:
            axis = self.axes
        elif isinstance(axis, str):
            axis = (axis,)
        axis = tuple(ax for ax in axis if ax in self.axes)
        idx = tuple(self.axes.index(ax) for ax in axis)
        axes = tuple(ax for ax in self.axes if ax not in axis)
        return ScalarArray(func(self.ndarray, axis=idx), axes)

    def min(self, axis=None):
        return self._reduce(np.min, axis)

    def max(self, axis=None):
        return self._reduce(np.max, axis)

    def sum(self, axis=None):
        return self._reduce(np.sum, axis)

    def mean(self, axis=None):
        return self._reduce(np.mean, axis)

    def ptp(self, axis=None):
        return self._reduce(np.ptp, axis)

    def cell_centers(self, axis) -> "ScalarArray":
        """Midpoints between adjacent vertices along `axis` (one name or several)."""
        result = self
        for ax in (axis,) if isinstance(axis, str) else tuple(axis):
            if ax not in result.axes:
                continue
            i = result.axes.index(ax)
            nd = np.moveaxis(result.ndarray, i, 0)
            nd = (nd[1:] + nd[:-1]) / 2
            result = ScalarArray(np.moveaxis(nd, 0, i), result.axes)
        return result

    def volume_cell(self, axis: str) -> "ScalarArray":
        """Signed length of the cells between adjacent vertices along `axis` (``na`` ``volume_cell``)."""
        return self[{axis: slice(1, None)}] - self[{axis: slice(None, -1)}]


def linspace(
    start,
    stop,
    axis: str,
    num: int,
    endpoint: bool = True,
    centers: bool = False,
) -> ScalarArray:
    """
    ``na.linspace``: `num` samples on a new axis.  With ``centers=True`` the
    samples are the centres of `num` equal cells spanning [start, stop]
    (the convention of the grids in ``optika/systems/_sequential.py:1916-1934``).
    """
    start = as_named_array(start)
    stop = as_named_array(stop)
    shp = shape_broadcasted(start, stop)
    a = aligned(start, shp)[..., None]
    b = aligned(stop, shp)[..., None]
    if centers:
        t = (np.arange(num) + 0.5) / num
    else:
        t = np.linspace(0, 1, num, endpoint=endpoint)
    nd = a + (b - a) * t
    nd = np.broadcast_to(nd, tuple(shp.values()) + (num,))
    return ScalarArray(nd, tuple(shp) + (axis,))


def arange(start, stop, axis: str, step=1) -> ScalarArray:
    return ScalarArray(np.arange(start, stop, step), (axis,))


def stack(arrays: Sequence, axis: str):
    """Stack scalars or vectors along a new leading axis named `axis`."""
    first = arrays[0]
    if isinstance(first, _VectorArray):
        return first._map_many(lambda *c: stack(list(c), axis), *arrays[1:])
    shp = shape_broadcasted(*arrays)
    nd = np.stack(
        [np.broadcast_to(aligned(a, shp), tuple(shp.values())) for a in arrays]
    )
    return ScalarArray(nd, (axis,) + tuple(shp))


class _VectorArray:
    """Component-wise container; arithmetic maps over the components."""

    __array_priority__ = 2000
    _names: tuple[str, ...] = ()

    @property
    def components(self) -> tuple:
        return tuple(getattr(self, n) for n in self._names)

    def _map(self, f):
        return type(self)(*[f(c) for c in self.components])

    def _map_many(self, f, *others):
        cols = [self.components] + [o.components for o in others]
        return type(self)(*[f(*c) for c in zip(*cols)])

    def _binary(self, other, f):
        if isinstance(other, _VectorArray):
            if type(other) is not type(self):
                return NotImplemented
            return type(self)(
                *[f(a, b) for a, b in zip(self.components, other.components)]
            )
        return self._map(lambda c: f(c, other))

    def __add__(self, o):
        return self._binary(o, lambda a, b: a + b)

    def __radd__(self, o):
        return self._binary(o, lambda a, b: b + a)

    def __sub__(self, o):
        return self._binary(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._binary(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._binary(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return self._binary(o, lambda a, b: b * a)

    def __truediv__(self, o):
        return self._binary(o, lambda a, b: a / b)

    def __neg__(self):
        return self._map(lambda c: -c)

    def __pos__(self):
        return self

    def __matmul__(self, o):
        if type(o) is not type(self):
            return NotImplemented
        result = 0
        for a, b in zip(self.components, o.components):
            result = result + a * b
        return result

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        # scalar (ufunc) vector: defer to the reflected vector operator
        return NotImplemented

    @property
    def shape(self) -> dict[str, int]:
        return shape_broadcasted(*self.components)

    @property
    def length(self):
        result = 0
        for c in self.components:
            result = result + np.square(c)
        return np.sqrt(result)

    @property
    def normalized(self):
        return self / self.length

    @property
    def value(self):
        return self

    @property
    def explicit(self):
        return self

    def broadcast_to(self, shape_: dict[str, int]):
        return self._map(lambda c: broadcast_to(c, shape_))

    def __getitem__(self, item):
        def index(c):
            c = as_named_array(c)
            return c[{ax: i for ax, i in item.items() if ax in c.axes}]

        return self._map(index)

    def min(self, axis=None):
        return self._map(lambda c: as_named_array(c).min(axis))

    def max(self, axis=None):
        return self._map(lambda c: as_named_array(c).max(axis))

    def ptp(self, axis=None):
        return self._map(lambda c: as_named_array(c).ptp(axis))

    def copy_shallow(self):
        return dataclasses.replace(self)

    def replace(self, **kwargs):
        return dataclasses.replace(self, **kwargs)


def _solid_angle_cells(components) -> np.ndarray:
    """
    Solid angles of the cells of a vertex grid of directions, given as three arrays broadcastable to
    ``[..., n_x + 1, n_y + 1]`` (not necessarily unit): ``[..., n_x, n_y]``.  Every cell is two spherical
    triangles (00, 10, 11) + (00, 11, 01), each ``2 atan2(a . (b x c), 1 + a.b + b.c + c.a)`` on the normalised
    corners.  The image of a 4096 x 4096 detector has 1.7e7 field cells: blocks of rows are normalised and worked on
    by a few threads (NumPy releases the GIL) instead of through two hundred 134 MB temporaries on one.
    """
    import concurrent.futures
    import os

    components = np.broadcast_arrays(*[np.asarray(c, dtype=np.float64) for c in components])
    shape_ = components[0].shape
    n_x, n_y = shape_[-2] - 1, shape_[-1] - 1
    out = np.empty(shape_[:-2] + (n_x, n_y))

    def dot(p, q):
        return p[0] * q[0] + p[1] * q[1] + p[2] * q[2]

    def triple(p, q, r):  # p . (q x r)
        return (p[0] * (q[1] * r[2] - q[2] * r[1]) + p[1] * (q[2] * r[0] - q[0] * r[2]) + p[2] * (q[0] * r[1] - q[1] * r[0]))

    def block(rows):
        i0, i1 = rows
        v = np.stack([c[..., i0:i1 + 1, :] for c in components])
        v /= np.sqrt(dot(v, v))
        a, b, c, d = v[..., :-1, :-1], v[..., 1:, :-1], v[..., 1:, 1:], v[..., :-1, 1:]
        ca = dot(c, a)
        first = np.arctan2(triple(a, b, c), 1 + dot(a, b) + dot(b, c) + ca)
        second = np.arctan2(triple(a, c, d), 1 + ca + dot(c, d) + dot(d, a))
        out[..., i0:i1, :] = 2 * first + 2 * second

    per_row = max(1, n_y * int(np.prod(shape_[:-2], dtype=np.int64)))
    rows = max(1, min(n_x, (1 << 20) // per_row))
    blocks = [(i, min(i + rows, n_x)) for i in range(0, n_x, rows)]
    workers = min(len(blocks), os.cpu_count() or 1, 16)
    if workers > 1:
        with concurrent.futures.ThreadPoolExecutor(workers) as pool:
            list(pool.map(block, blocks))
    else:
        for r in blocks:
            block(r)
    return out


def _corners(v: "_VectorArray", axis: tuple[str, str]):
    """The four corner vertices (00, 10, 11, 01) of every cell of a 2-D vertex grid."""
    ax, ay = axis
    shape_ = v.shape
    if ax not in shape_ or ay not in shape_:
        raise ValueError(f"axes {axis} must both be present in {shape_}")
    v = v.broadcast_to(shape_)
    lo, hi = slice(None, -1), slice(1, None)
    return v[{ax: lo, ay: lo}], v[{ax: hi, ay: lo}], v[{ax: hi, ay: hi}], v[{ax: lo, ay: hi}]


@dataclasses.dataclass(eq=False)
class Cartesian2dVectorArray(_VectorArray):
    x: float | ScalarArray = 0
    y: float | ScalarArray = 0
    _names = ("x", "y")

    def volume_cell(self, axis: tuple[str, str]) -> ScalarArray:
        """Signed area of every vertex quadrilateral: half the cross product of its diagonals."""
        v00, v10, v11, v01 = _corners(self, axis)
        d1, d2 = v11 - v00, v01 - v10
        return (d1.x * d2.y - d1.y * d2.x) / 2


@dataclasses.dataclass(eq=False)
class Cartesian3dVectorArray(_VectorArray):
    x: float | ScalarArray = 0
    y: float | ScalarArray = 0
    z: float | ScalarArray = 0
    _names = ("x", "y", "z")

    @property
    def xy(self) -> Cartesian2dVectorArray:
        return Cartesian2dVectorArray(self.x, self.y)

    def solid_angle_cell(self, axis: tuple[str, str]) -> ScalarArray:
        """
        Signed solid angle [sr] of the spherical quadrilateral spanned by the four direction
        vertices of every cell: two spherical triangles, each by Van Oosterom & Strackee (1983).
        """
        ax, ay = axis
        shape_ = self.shape
        if ax not in shape_ or ay not in shape_:
            raise ValueError(f"axes {axis} must both be present in {shape_}")
        other = tuple(a for a in shape_ if a not in axis)
        order = other + (ax, ay)
        full = {a: shape_[a] for a in order}
        out = _solid_angle_cells([aligned(as_named_array(getattr(self, c)), full) for c in self._names])
        return ScalarArray(out, order)

    def cross(self, o: "Cartesian3dVectorArray") -> "Cartesian3dVectorArray":
        return Cartesian3dVectorArray(
            x=self.y * o.z - self.z * o.y,
            y=self.z * o.x - self.x * o.z,
            z=self.x * o.y - self.y * o.x,
        )


def Cartesian2dVectorLinearSpace(
    start,
    stop,
    axis: Cartesian2dVectorArray,
    num,
    endpoint: bool = True,
    centers: bool = False,
) -> Cartesian2dVectorArray:
    """
    ``na.Cartesian2dVectorLinearSpace``: `x` varies along ``axis.x`` only and `y`
    along ``axis.y`` only (a separable grid), as used for the field and pupil
    grids in ``optika/systems/_sequential.py:1916-1934``.
    """

    def comp(v, name):
        return getattr(v, name) if isinstance(v, Cartesian2dVectorArray) else v

    return Cartesian2dVectorArray(
        x=linspace(
            comp(start, "x"), comp(stop, "x"), axis.x, comp(num, "x"), endpoint, centers
        ),
        y=linspace(
            comp(start, "y"), comp(stop, "y"), axis.y, comp(num, "y"), endpoint, centers
        ),
    )


@dataclasses.dataclass(eq=False)
class FunctionArray:
    """``na.FunctionArray``: a pair of (inputs, outputs) defined on the same grid."""

    inputs: object = None
    outputs: object = None

    def copy_shallow(self):
        return dataclasses.replace(self)
