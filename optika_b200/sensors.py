"""
Imaging sensors.

Mirrors ``optika.sensors.ImagingSensor`` (``optika/sensors/_sensors.py:38-171,
431-500``): as a surface it is a flat sag with a rectangular aperture of
``width_pixel * num_pixel / 2`` (``:83-90``); ``collect`` bins a ray cloud onto
the pixel grid (``:92-171``).  The binning runs on the device
(``optk_bin`` / the fused ``optk_trace`` image path): warp-aggregated fp64
reductions into the detector planes.
"""

from __future__ import annotations
from typing import Sequence
import ctypes as C
import dataclasses
import numpy as np
from . import named as na
from . import units as u
from . import sags as _sags
from . import apertures as _apertures
from . import materials as _materials
from . import _lib as L
from .surfaces import AbstractSurface
from .transformations import AbstractTransformation
from .vectors import SpectralPositionalVectorArray

from ._detector_physics import (  # noqa: E402,F401  (optika.sensors.charge_diffusion, electrons_measured, ...)
    charge_diffusion, mean_charge_capture, kernel_diffusion, energy_bandgap, energy_pair, energy_pair_inf,
    quantum_yield_ideal, fano_factor, fano_factor_inf, probability_of_n_pairs, electrons_measured,
)

__all__ = [
    "IdealSensorMaterial", "AbstractImagingSensor", "ImagingSensor",
    "charge_diffusion", "mean_charge_capture", "kernel_diffusion", "energy_bandgap", "energy_pair", "energy_pair_inf",
    "quantum_yield_ideal", "fano_factor", "fano_factor_inf", "probability_of_n_pairs", "electrons_measured",
]


_NOISE_BLOCK = 1 << 20


def _noise_blocks(draw, out: np.ndarray, seed, stream: int) -> np.ndarray:
    """
    Fill the flat array `out` with ``draw(rng, slice)`` block by block on a few threads.  Every block of 2^20 pixels
    has its own generator, spawned from (seed, stream, block index): the result depends on the seed and on nothing
    else (not on the number of threads), and a 4096 x 4096 x 8 exposure takes 0.5 s instead of 4.5 s.
    """
    import concurrent.futures
    import os

    n = out.size
    blocks = [(k, slice(i, min(i + _NOISE_BLOCK, n))) for k, i in enumerate(range(0, n, _NOISE_BLOCK))]
    entropy = 0 if seed is None else int(seed)
    if seed is None:
        entropy = int(np.random.SeedSequence().entropy)

    def work(block):
        k, where = block
        rng = np.random.default_rng(np.random.SeedSequence(entropy, spawn_key=(stream, k)))
        out[where] = draw(rng, where)

    workers = min(len(blocks), os.cpu_count() or 1, 16)
    if workers > 1:
        with concurrent.futures.ThreadPoolExecutor(workers) as pool:
            list(pool.map(work, blocks))
    else:
        for b in blocks:
            work(b)
    return out


def shot_noise(photons: np.ndarray, seed) -> np.ndarray:
    """Poisson-distributed counts with the given expectations (negative expectations count as zero)."""
    lam = np.maximum(np.ascontiguousarray(photons, dtype=np.float64), 0).reshape(-1)
    out = np.empty(lam.size, dtype=np.int64)
    return _noise_blocks(lambda rng, where: rng.poisson(lam[where]), out, seed, 0).reshape(np.shape(photons))


def read_noise(electrons: np.ndarray, sigma: float, seed) -> np.ndarray:
    """`electrons` plus zero-mean Gaussian noise of `sigma` electrons."""
    loc = np.ascontiguousarray(electrons, dtype=np.float64).reshape(-1)
    out = np.empty(loc.size, dtype=np.float64)
    return _noise_blocks(lambda rng, where: rng.normal(loc=loc[where], scale=sigma), out, seed, 1).reshape(np.shape(electrons))


@dataclasses.dataclass(eq=False)
class IdealSensorMaterial(_materials.Vacuum):
    """
    Unit quantum efficiency, no charge diffusion
    (``optika/sensors/materials/_materials.py:1566-1643``):
    ``direction_refracted = -direction . normal`` and photons map 1:1 to electrons.
    """

    uses_direction = False  # nothing below depends on the angle of incidence: callers need not form it

    def signal(self, photons, wavelength=None, direction=1, noise: bool = False, rng=None, **kwargs):
        # optika/sensors/materials/_materials.py:1576-1601; shot noise is drawn on the host from seeded NumPy
        # generators, one per block of pixels (`shot_noise`); `rng` is the seed (or a Generator to take one from)
        if noise:
            if isinstance(rng, np.random.Generator):
                rng = int(rng.integers(0, 2**63 - 1))
            p = na.as_named_array(photons)
            photons = na.ScalarArray(shot_noise(p.ndarray, rng), p.axes)
        return photons


@dataclasses.dataclass(eq=False)
class AbstractImagingSensor(AbstractSurface):
    @property
    def sag(self) -> _sags.AbstractSag:
        return _sags.NoSag()  # _sensors.py:45-46

    @property
    def rulings(self) -> None:
        return None  # _sensors.py:48-50

    @property
    def width_pixel_xy(self) -> na.Cartesian2dVectorArray:
        w = self.width_pixel
        if isinstance(w, na.Cartesian2dVectorArray):
            return na.Cartesian2dVectorArray(u.length(w.x), u.length(w.y))
        w = u.length(w)
        return na.Cartesian2dVectorArray(w, w)

    @property
    def aperture(self) -> _apertures.RectangularAperture:
        """The light-sensitive area (``_sensors.py:83-90``)."""
        w = self.width_pixel_xy
        return _apertures.RectangularAperture(
            half_width=na.Cartesian2dVectorArray(
                w.x * self.num_pixel.x / 2,
                w.y * self.num_pixel.y / 2,
            ),
        )

    def pixel_edges(self) -> tuple[np.ndarray, np.ndarray]:
        """
        ``linspace(bound_lower, bound_upper, num_pixel + 1)`` (``_sensors.py:141-149``),
        with the bounds taken from the aperture's vertices as the reference does.
        """
        ap = self.aperture
        lo, hi = ap.bound_lower, ap.bound_upper
        ex = np.linspace(float(lo.x), float(hi.x), int(self.num_pixel.x) + 1)
        ey = np.linspace(float(lo.y), float(hi.y), int(self.num_pixel.y) + 1)
        return ex, ey

    def collect(
        self,
        rays,
        wavelength: na.ScalarArray,
        axis: None | str | Sequence[str] = None,
        where=True,
        device=None,
    ):
        """
        Bin rays given in sensor-local coordinates onto the pixel grid
        (``optika/sensors/_sensors.py:92-171``).  Returns ``(image, direction)``:
        `image` is a :class:`~optika_b200.named.FunctionArray` whose inputs are
        the bin edges and whose outputs are the binned intensity; `direction` the
        flux-weighted mean cosine of the refracted angle (1 where empty).

        `wavelength` holds the bin EDGES on one named axis; `axis` lists the
        logical axes to sum over (default: all axes of `rays`).
        """
        from . import _engine

        torch = _engine._torch()
        device = _engine.require_cuda(device)
        if not isinstance(rays, _engine.DeviceRays):
            rays = _host_to_device(rays, device)
        shape_ = rays.shape
        if axis is None:
            axis = tuple(shape_)
        elif isinstance(axis, str):
            axis = (axis,)
        axis = tuple(ax for ax in axis if ax in shape_)
        keep = [ax for ax in shape_ if ax not in axis]
        order = keep + list(axis)
        perm = [list(shape_).index(ax) for ax in order]
        dims = [shape_[ax] for ax in order]
        n_keep = int(np.prod([shape_[ax] for ax in keep], dtype=np.int64)) if keep else 1
        n_bin = rays.size // max(n_keep, 1)

        def arrange(t):
            t = t.reshape(tuple(shape_.values()))
            if perm != list(range(len(perm))):
                t = t.permute(perm).contiguous()
            return t.reshape(n_keep, n_bin)

        w = arrange(rays.fields["wavelength"])
        x = arrange(rays.fields["px"])
        y = arrange(rays.fields["py"])
        dz = arrange(rays.fields["dz"])
        inten = arrange(rays.fields["intensity"])
        mask = arrange(rays.unvignetted)
        if where is not True:
            wh = na.broadcast_to(na.as_named_array(where), shape_)
            wh = torch.from_numpy(np.ascontiguousarray(wh.ndarray).astype(np.uint8)).to(device)
            mask = mask & arrange(wh)

        wavelength = na.as_named_array(u.length(wavelength))
        if wavelength.ndim != 1:
            raise ValueError("`wavelength` must hold the bin edges along exactly one axis")
        (axis_wavelength,) = wavelength.axes
        ex, ey = self.pixel_edges()
        image = _engine.DeviceImage.zeros(
            wavelength.ndarray, ex, ey, device, leading=(n_keep,), moments=True, counts=False
        )
        stream = _engine._stream_ptr(device)
        for k in range(n_keep):
            im = image.struct(k)
            L.check(
                L.lib().optk_bin(
                    n_bin, w[k].data_ptr(), x[k].data_ptr(), y[k].data_ptr(), dz[k].data_ptr(),
                    inten[k].data_ptr(), mask[k].data_ptr(), C.byref(im), stream,
                )
            )
        flux = image.flux.cpu().numpy()
        moment = image.moment_real.cpu().numpy()
        keep_dims = [shape_[ax] for ax in keep]
        axes_out = tuple(keep) + (axis_wavelength, self.axis_pixel.x, self.axis_pixel.y)
        flux = flux.reshape(keep_dims + list(flux.shape[1:]))
        moment = moment.reshape(flux.shape)
        nonempty = flux > 0
        with np.errstate(invalid="ignore", divide="ignore"):
            direction = np.where(nonempty, moment / flux, 1) + 0j  # _sensors.py:163-169
        inputs = SpectralPositionalVectorArray(
            wavelength=wavelength,
            position=na.Cartesian2dVectorArray(
                x=na.ScalarArray(ex, (self.axis_pixel.x,)),
                y=na.ScalarArray(ey, (self.axis_pixel.y,)),
            ),
        )
        return (
            na.FunctionArray(inputs=inputs, outputs=na.ScalarArray(flux, axes_out)),
            na.ScalarArray(direction, axes_out),
        )

    def expose(
        self, image: na.FunctionArray, direction=1, axis_wavelength=None, timedelta=None, noise: bool = False,
        seed=None,
    ):
        """
        Photons -> electrons (``_sensors.py:173-252``).  `noise` adds Poisson shot noise
        (in the material) and zero-mean Gaussian read noise of ``read_noise`` electrons
        (``:243-248``), drawn on the host from NumPy generators spawned from `seed`, one per block of 2^20 pixels
        (`shot_noise`, `read_noise`).  `direction` may be a callable that forms it on demand.
        """
        if timedelta is None:
            timedelta = self.timedelta_exposure
        photons = image.outputs * timedelta
        if noise and seed is None:
            seed = int(np.random.SeedSequence().entropy % (2**63))
        if callable(direction):  # formed only for materials that look at it (a full-size complex array otherwise)
            direction = direction() if getattr(self.material, "uses_direction", True) else 1
        electrons = self.material.signal(photons=photons, direction=direction, noise=noise, rng=seed if noise else None)
        if noise:
            e = na.as_named_array(electrons)
            electrons = na.ScalarArray(read_noise(e.ndarray, float(self.read_noise), seed), e.axes)
        return dataclasses.replace(image, outputs=electrons)

    def measure(
        self, rays, wavelength, axis=None, where=True, axis_wavelength=None, timedelta=None, noise: bool = False,
        seed=None,
    ):
        """``collect`` then ``expose`` (``_sensors.py:374-428``)."""
        image, direction = self.collect(rays, wavelength, axis=axis, where=where)
        return self.expose(
            image, direction, axis_wavelength=axis_wavelength, timedelta=timedelta, noise=noise, seed=seed
        )


def _host_to_device(rays, device):
    from . import _engine

    torch = _engine._torch()
    shape_ = rays.shape
    dims = tuple(shape_.values())

    def dev(v, dtype):
        nd = np.ascontiguousarray(na.broadcast_to(na.as_named_array(v), shape_).ndarray.astype(dtype))
        return torch.from_numpy(nd).to(device).reshape(-1)

    fields = {name: dev(get(rays), np.float64) for name, get in _engine._FIELD_GETTERS}
    mask = dev(rays.unvignetted, np.uint8)
    return _engine.DeviceRays(fields, mask, dict(shape_))


@dataclasses.dataclass(eq=False)
class ImagingSensor(AbstractImagingSensor):
    """An imaging sensor (``optika/sensors/_sensors.py:431-500``)."""

    name: None | str = None
    width_pixel: float | na.Cartesian2dVectorArray = 0
    axis_pixel: na.Cartesian2dVectorArray = None
    num_pixel: na.Cartesian2dVectorArray = None
    timedelta_exposure: float = 1.0
    read_noise: float = 0.0
    material: object = None
    aperture_mechanical: object = None
    is_field_stop: bool = False
    is_pupil_stop: bool = False
    transformation: None | AbstractTransformation = None
    kwargs_plot: None | dict = None

    def __post_init__(self):
        if self.material is None:
            self.material = IdealSensorMaterial()

    @property
    def shape(self) -> dict[str, int]:
        return na.broadcast_shapes(na.shape(self.transformation))
