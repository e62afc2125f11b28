"""
TEST / BASELINE INFRASTRUCTURE -- never imported by the product.

The one native kernel on the reference's path: the vector form of Snell's law compiled by numba as
a ``guvectorize`` ufunc with ``target="parallel"`` (``optika/materials/_snells_law.py:294-366``).
BASELINE.md section 4 asks the CPU arm to run that step the way the reference does, so the
oracle's NumPy expression (``oracle/raytrace.py::snells_law``, which follows the same lines
``:341-366``) gets a numba twin here, in the reference's two flavours of use:

* :func:`snells_law_parallel` -- ``target="parallel"``: numba's own thread pool over the whole
  array (how the reference calls it from single-threaded NumPy code);
* :func:`snells_law_serial` -- ``target="cpu"``: for a CPU arm that already splits the rays over
  host threads itself.

Both release the GIL and agree with the NumPy expression to the last bit (same operations in the
same order, no fast-math).
"""

from __future__ import annotations
import math

import numba as nb

_SIGNATURE = [
    "void(float64,float64,float64,float64,float64,float64,float64,float64,boolean,float64[:],float64[:],float64[:])"
]
_LAYOUT = "(),(),(),(),(),(),(),(),()->(),(),()"


def _body(ax, ay, az, n1, n2, ux, uy, uz, mirror, bx, by, bz):  # pragma: no cover (compiled)
    # b = r (a + d u),  r = n1 / n2,  d = -(a.u) - sign(a.u) (2 mirror - 1) sqrt(1 / r^2 + (a.u)^2 - |a|^2)
    squared_length = ax * ax + ay * ay + az * az
    ratio = n1 / n2
    projection = ax * ux + ay * uy + az * uz
    flip = -math.copysign(1.0, projection)
    d = -projection + flip * (2 * mirror - 1) * math.sqrt(1 / (ratio * ratio) + projection * projection - squared_length)
    bx[0] = ratio * (ax + d * ux)
    by[0] = ratio * (ay + d * uy)
    bz[0] = ratio * (az + d * uz)


snells_law_parallel = nb.guvectorize(_SIGNATURE, _LAYOUT, target="parallel", nopython=True, cache=False)(_body)
snells_law_serial = nb.guvectorize(_SIGNATURE, _LAYOUT, target="cpu", nopython=True, cache=False)(_body)


def threads() -> int:
    return nb.get_num_threads()
