"""
A NumPy-oracle backend for ``optika_b200._stops`` (test infrastructure): the same
stop solver, with every trace and sag evaluation done by ``oracle/raytrace.py``.
"""

import numpy as np

from optika_b200 import named as na
from optika_b200.rays import RayVectorArray
from oracle import raytrace as ora

import configs


class OracleBackend:
    @staticmethod
    def propagate(surfaces, rays):
        r0, shape_ = configs.flatten_rays(rays)
        dims = tuple(shape_.values())
        out = ora.propagate_rays(surfaces, {k: v.reshape(-1) for k, v in r0.items()}, converge=True, extended=True)
        axes = tuple(shape_)

        def get(name, dtype=float):
            return na.ScalarArray(out[name].reshape(dims).astype(dtype), axes)

        return RayVectorArray(
            wavelength=get("wavelength"),
            position=na.Cartesian3dVectorArray(get("px"), get("py"), get("pz")),
            direction=na.Cartesian3dVectorArray(get("dx"), get("dy"), get("dz")),
            intensity=get("intensity"),
            attenuation=get("attenuation"),
            index_refraction=get("index_refraction"),
            unvignetted=get("unvignetted", bool),
        )

    @staticmethod
    def sag(surface, x, y):
        shape_ = na.shape_broadcasted(x, y)
        xx = np.broadcast_to(na.aligned(na.as_named_array(x), shape_), tuple(shape_.values()))
        yy = np.broadcast_to(na.aligned(na.as_named_array(y), shape_), tuple(shape_.values()))
        return na.ScalarArray(ora.sag_value(surface.sag, xx, yy), tuple(shape_))
