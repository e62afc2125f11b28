"""
Host side of the device engine: flattening named ray grids into strided
structure-of-arrays buffers, launching ``liboptk`` through ctypes, and wrapping
the results back into named arrays.

This module is the drop-in boundary of SURVEY.md section 8b: everything above it
keeps the reference's Python signatures, everything below it is the C ABI of
``include/optk.h``.  PyTorch is used for device memory, streams and (in
:mod:`optika_b200.distributed`) NCCL only; all arithmetic on rays happens in the
CUDA kernels of ``optika_b200/csrc``.
"""

from __future__ import annotations
import ctypes as C
import dataclasses
import math
import numpy as np
from typing import Sequence
from . import named as na
from . import units as u
from . import _lib as L
from . import _lowering
from .rays import RayVectorArray

__all__ = [
    "CompiledSystem",
    "DeviceRays",
    "DeviceImage",
    "trace",
    "require_cuda",
]

_MAX_LAUNCH = 2**31 - 1


def _torch():
    import torch

    return torch


def require_cuda(device=None):
    """The torch CUDA device to run on; raises when there is none (no CPU fallback)."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise L.OptkError(
            "optika_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _stream_ptr(device) -> int:
    return _torch().cuda.current_stream(device).cuda_stream


@dataclasses.dataclass(eq=False)
class DeviceRays:
    """
    Rays resident in HBM: one dense fp64 tensor per field (structure of arrays)
    plus the uint8 ``unvignetted`` mask, all of the named `shape` (C order).
    """

    fields: dict
    unvignetted: object
    shape: dict[str, int]
    cos_incidence: object = None  # -a . n at the last traced surface, when captured

    def last_state(self, axis_position: int | None = None) -> "DeviceRays":
        """The state after the last surface of an accumulated trace (`[config][surface][rays...]`)."""
        axes = list(self.shape)
        k = self._surface_axis
        n_lead = int(np.prod([self.shape[a] for a in axes[:k]], dtype=np.int64)) if k else 1
        n_s = self.shape[axes[k]]

        def pick(t):
            return t.reshape(n_lead, n_s, -1)[:, -1].contiguous().reshape(-1)

        shape_ = {a: n for a, n in self.shape.items() if a != axes[k]}
        return DeviceRays({f: pick(t) for f, t in self.fields.items()}, pick(self.unvignetted), shape_)

    _surface_axis = 0

    @classmethod
    def concatenate(cls, parts: list, axis: str) -> "DeviceRays":
        """Join accumulated traces of consecutive surface ranges along the surface axis."""
        torch = _torch()
        first = parts[0]
        axes = list(first.shape)
        k = axes.index(axis)
        n_lead = int(np.prod([first.shape[a] for a in axes[:k]], dtype=np.int64)) if k else 1

        def join(get):
            return torch.cat([get(p).reshape(n_lead, p.shape[axis], -1) for p in parts], dim=1).reshape(-1)

        shape_ = dict(first.shape)
        shape_[axis] = sum(p.shape[axis] for p in parts)
        out = cls({f: join(lambda p, f=f: p.fields[f]) for f in first.fields}, join(lambda p: p.unvignetted), shape_)
        out._surface_axis = k
        return out

    @property
    def size(self) -> int:
        return int(np.prod(list(self.shape.values()), dtype=np.int64)) if self.shape else 1

    def to_host(self) -> RayVectorArray:
        axes = tuple(self.shape)
        dims = tuple(self.shape.values())

        def get(name):
            return na.ScalarArray(self.fields[name].reshape(dims).cpu().numpy(), axes)

        return RayVectorArray(
            wavelength=get("wavelength"),
            position=na.Cartesian3dVectorArray(get("px"), get("py"), get("pz")),
            direction=na.Cartesian3dVectorArray(get("dx"), get("dy"), get("dz")),
            intensity=get("intensity"),
            attenuation=get("attenuation"),
            index_refraction=get("index_refraction"),
            unvignetted=na.ScalarArray(
                self.unvignetted.reshape(dims).cpu().numpy().astype(bool), axes
            ),
        )

    def __getitem__(self, item: dict) -> "DeviceRays":
        axes = tuple(self.shape)
        dims = tuple(self.shape.values())
        index = tuple(item.get(ax, slice(None)) for ax in axes)
        new_axes = [ax for ax in axes if not (ax in item and not isinstance(item[ax], slice))]

        def get(t):
            return t.reshape(dims)[index].contiguous()

        fields = {k: get(v) for k, v in self.fields.items()}
        mask = get(self.unvignetted)
        shape_ = dict(zip(new_axes, fields["wavelength"].shape))
        return DeviceRays(fields, mask, shape_)


def _is_linspace(edges) -> bool:
    """Every edge within 1e-9 of a bin width of ``first + i (last - first) / n`` (``optk_image_t.uniform_edges``)."""
    e = np.asarray(edges, dtype=np.float64)
    n = len(e) - 1
    if n < 1 or not np.all(np.isfinite(e)) or not e[-1] > e[0]:
        return False
    width = (e[-1] - e[0]) / n
    return bool(np.max(np.abs(e - (e[0] + np.arange(n + 1) * width))) <= 1e-9 * width)


@dataclasses.dataclass(eq=False)
class DeviceImage:
    """Detector planes in HBM, ``[n_wavelength][n_x][n_y]`` (optionally with leading config axes)."""

    edges_wavelength: object
    edges_x: object
    edges_y: object
    flux: object
    moment_real: object = None
    moment_imag: object = None
    counts: object = None
    range: object = None  # first / last edge of (wavelength, x, y) as host floats, when known
    uniform_edges: int = 0  # optk_image_t.uniform_edges: bit 0 / 1 -- the x / y edges are a linspace

    buffer_f64: object = None  # fused layout: [n_config][row] with flux | moment_real of one configuration side by side
    buffer_i64: object = None  # fused layout: [n_config][row] counts

    @classmethod
    def zeros(cls, edges_wavelength, edges_x, edges_y, device, leading=(), moments=True, counts=True,
              fused: bool = False, pad_to: int = 1):
        """
        Zeroed planes.  ``fused``: the fp64 planes of one configuration are adjacent in ONE buffer
        (``buffer_f64[c] = flux[c] | moment_real[c]``, rows padded to a multiple of `pad_to` elements;
        ``buffer_i64[c] = counts[c]``), so a configuration is reduced over the ranks with one collective
        per dtype and read back with one copy (:class:`optika_b200.distributed.ImagePipeline`).
        """
        torch = _torch()

        def dev(e):
            return torch.from_numpy(np.array(e, dtype=np.float64)).to(device)

        ew, ex, ey = dev(edges_wavelength), dev(edges_x), dev(edges_y)
        dims3 = (len(ew) - 1, len(ex) - 1, len(ey) - 1)
        dims = tuple(leading) + dims3
        rng = [float(v) for e in (edges_wavelength, edges_x, edges_y) for v in (np.asarray(e)[0], np.asarray(e)[-1])]
        uniform = (1 if _is_linspace(edges_x) else 0) | (2 if _is_linspace(edges_y) else 0)
        if fused:
            n = int(np.prod(dims3, dtype=np.int64))
            n_config = int(np.prod(tuple(leading), dtype=np.int64)) if leading else 1
            n_planes = 2 if moments else 1
            pad = lambda m: -(-m // pad_to) * pad_to  # noqa: E731
            buf_f = torch.zeros((n_config, pad(n_planes * n)), dtype=torch.float64, device=device)
            buf_i = torch.zeros((n_config, pad(n)), dtype=torch.int64, device=device) if counts else None
            plane = lambda b, k: b[:, k * n:(k + 1) * n].view(dims)  # noqa: E731
            return cls(
                ew, ex, ey, flux=plane(buf_f, 0), moment_real=plane(buf_f, 1) if moments else None, moment_imag=None,
                counts=plane(buf_i, 0) if counts else None, range=rng, buffer_f64=buf_f, buffer_i64=buf_i,
                uniform_edges=uniform,
            )
        z = lambda dt: torch.zeros(dims, dtype=dt, device=device)  # noqa: E731
        return cls(
            ew, ex, ey,
            flux=z(torch.float64),
            moment_real=z(torch.float64) if moments else None,
            moment_imag=None,
            counts=z(torch.int64) if counts else None,
            range=rng,
            uniform_edges=uniform,
        )

    def zero_(self):
        """Clear every plane (re-use of the buffers between exposures)."""
        if self.buffer_f64 is not None:
            self.buffer_f64.zero_()
            if self.buffer_i64 is not None:
                self.buffer_i64.zero_()
        else:
            for t in (self.flux, self.moment_real, self.moment_imag, self.counts):
                if t is not None:
                    t.zero_()
        return self

    _pinned = {}

    def to_host(self, pinned: bool = False) -> dict:
        """
        Copy the planes to host memory; returns ``{name: torch tensor}``.  With ``pinned`` the
        destination buffers are page-locked and cached per (name, shape) -- the copy then runs at
        PCIe speed instead of being staged through pageable memory, but the SAME buffers are handed
        out again by the next call with equal shapes (callers that keep two images must copy);
        :meth:`release_pinned` frees the cache.  The default returns fresh pageable tensors.
        """
        torch = _torch()
        out = {}
        for name in ("flux", "moment_real", "moment_imag", "counts"):
            t = getattr(self, name)
            if t is None:
                continue
            if pinned:
                key = (name, tuple(t.shape), t.dtype)
                buf = DeviceImage._pinned.get(key)
                if buf is None:
                    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    DeviceImage._pinned[key] = buf
                buf.copy_(t, non_blocking=True)
                out[name] = buf
            else:
                out[name] = t.cpu()
        if pinned:
            torch.cuda.current_stream(self.flux.device).synchronize()
        return out

    @classmethod
    def release_pinned(cls):
        """Free the page-locked read-back buffers cached by ``to_host(pinned=True)``."""
        cls._pinned.clear()

    def struct(self, plane_index: int = 0) -> L.Image:
        im = L.Image()
        im.n_wavelength = len(self.edges_wavelength) - 1
        im.n_x = len(self.edges_x) - 1
        im.n_y = len(self.edges_y) - 1
        im.edges_wavelength = self.edges_wavelength.data_ptr()
        im.edges_x = self.edges_x.data_ptr()
        im.edges_y = self.edges_y.data_ptr()

        def ptr(t):
            # planes are [config axes...][n_w][n_x][n_y]; the configuration axes are contiguous among
            # themselves in both layouts, so the last of them carries the stride of one configuration
            if t is None:
                return None
            return t.data_ptr() + (plane_index * t.stride(t.dim() - 4) * 8 if t.dim() > 3 else 0)

        if self.range is not None:
            im.has_range = 1
            im.range[:] = self.range
            im.uniform_edges = self.uniform_edges
        im.flux = ptr(self.flux)
        im.moment_real = ptr(self.moment_real)
        im.moment_imag = ptr(self.moment_imag)
        im.counts = ptr(self.counts)
        return im


@dataclasses.dataclass(eq=False)
class DeviceGroups:
    """
    Per-group accumulators in HBM for the reductions fused into the trace (``optk_image_t.group_size``):
    one entry per group of `group_size` consecutive rays of a configuration -- with the pupil axes
    innermost, the pupil of one field point.  ``flux`` = sum of intensity, ``moment_real`` / ``moment_imag``
    = sums of x / y, ``counts`` = number, over the unvignetted rays; shape ``[n_config][n_groups]``.
    """

    group_size: int
    n_groups: int
    flux: object
    moment_real: object
    moment_imag: object
    counts: object

    @classmethod
    def zeros(cls, n_config: int, n_groups: int, group_size: int, device):
        torch = _torch()
        z = lambda dt: torch.zeros((n_config, n_groups), dtype=dt, device=device)  # noqa: E731
        return cls(group_size, n_groups, z(torch.float64), z(torch.float64), z(torch.float64), z(torch.int64))

    def struct(self, plane_index: int = 0, first_ray: int = 0) -> L.Image:
        """The accumulators of configuration `plane_index` for a launch that starts at ray `first_ray`."""
        if first_ray % self.group_size:
            raise ValueError("a launch must start at a group boundary")
        first = first_ray // self.group_size
        im = L.Image()
        im.n_wavelength, im.n_x, im.n_y = 1, self.n_groups - first, 1
        im.group_size = self.group_size
        offset = 8 * (plane_index * self.n_groups + first)
        im.flux = self.flux.data_ptr() + offset
        im.moment_real = self.moment_real.data_ptr() + offset
        im.moment_imag = self.moment_imag.data_ptr() + offset
        im.counts = self.counts.data_ptr() + offset
        return im


class CompiledSystem:
    """
    A surface list lowered to the device table (``optk_system_create``): the
    replacement for iterating ``SequentialSystem.surfaces_all`` in Python
    (``optika/propagators.py:38-39, 67-71``).
    """

    def __init__(self, surfaces, stages: int = L.STAGE_ALL, local_last: bool = False, lowered=None):
        self.surfaces = list(surfaces)
        table, shape_ = lowered if lowered is not None else _lowering.lower_system(self.surfaces, stages=stages)
        # surfaces whose efficiency is a per-ray multilayer evaluation (MultilayerMirror /
        # MultilayerFilm): the trace is chained around them, see `trace`
        self.coatings = {
            k: s.material for k, s in enumerate(self.surfaces)
            if stages == L.STAGE_ALL and hasattr(s.material, "efficiency_device")
        }
        if local_last:
            # leave the rays in the local frame of the last surface
            n = len(self.surfaces)
            for k in range(n - 1, len(table), n):
                table[k].flags |= L.F_LOCAL_OUT
        self.table = table
        self.shape = shape_
        self.n_surface = len(self.surfaces)
        self.n_config = len(table) // max(self.n_surface, 1)
        self._handles = {}  # device ordinal -> optk_system_t*: efficiency tables live on ONE device
        self.handle  # validate the table now (unsupported kinds raise here, as before)

    def handle_for(self, device=None):
        """The ``optk_system_t*`` whose device-side tables live on `device` (default: the current one)."""
        torch = _torch()
        if not torch.cuda.is_available():
            ordinal = -1
        elif device is None:
            ordinal = torch.cuda.current_device()
        else:
            ordinal = torch.device(device).index
            ordinal = torch.cuda.current_device() if ordinal is None else ordinal
        handle = self._handles.get(ordinal)
        if handle is None:
            handle = C.c_void_p()
            with device_guard(ordinal):
                L.check(L.lib().optk_system_create(self.table, self.n_surface, self.n_config, C.byref(handle)))
            self._handles[ordinal] = handle
        return handle

    @property
    def handle(self):
        return self.handle_for(None)

    def __del__(self):
        try:
            for handle in getattr(self, "_handles", {}).values():
                L.lib().optk_system_destroy(handle)
            self._handles = {}
        except Exception:  # pragma: no cover
            pass


def device_guard(device):
    """
    Context manager that makes `device` the current CUDA device (no-op without CUDA or for ``-1``).
    Launches themselves follow the stream they are given (``DeviceScope`` in ``csrc/api.cu``); the
    guard is for what has no stream argument: allocations made by ``optk_system_create`` and
    torch's own current-stream lookups.
    """
    import contextlib

    torch = _torch()
    if not torch.cuda.is_available() or device is None or device == -1:
        return contextlib.nullcontext()
    return torch.cuda.device(device)


# ---------------------------------------------------------------------------
# flattening
# ---------------------------------------------------------------------------
_FIELD_GETTERS = (
    ("wavelength", lambda r: r.wavelength),
    ("px", lambda r: r.position.x),
    ("py", lambda r: r.position.y),
    ("pz", lambda r: r.position.z),
    ("dx", lambda r: r.direction.x),
    ("dy", lambda r: r.direction.y),
    ("dz", lambda r: r.direction.z),
    ("intensity", lambda r: r.intensity),
    ("attenuation", lambda r: r.attenuation),
    ("index_refraction", lambda r: r.index_refraction),
)


class _View:
    """A device tensor plus its element stride along every axis of the full grid."""

    def __init__(self, tensor, strides: dict[str, int]):
        self.tensor = tensor
        self.strides = strides


def _view_of(value, device, is_mask=False) -> _View:
    """Upload a (small, possibly broadcast) host value; strides are per named axis."""
    torch = _torch()
    dtype = np.uint8 if is_mask else np.float64
    if isinstance(value, na.ScalarArray):
        nd = np.ascontiguousarray(np.asarray(value.ndarray).astype(dtype, copy=False))
        strides = {}
        for ax, st, n in zip(value.axes, nd.strides, nd.shape):
            strides[ax] = 0 if n == 1 else st // nd.itemsize
        t = torch.from_numpy(nd.reshape(-1).copy()).to(device)
        return _View(t, strides)
    nd = np.asarray(u.length(value)).astype(dtype).reshape(1)
    return _View(torch.from_numpy(nd).to(device), {})


def _grid_shape(rays, config_shape_: dict[str, int], normal=None) -> tuple[dict[str, int], dict[str, int]]:
    """(configuration axes, ray axes) of the full grid; configuration axes come first."""
    rays_shape = na.broadcast_shapes(rays.shape, na.shape(normal))
    total = na.broadcast_shapes(config_shape_, rays_shape)
    config = {ax: total[ax] for ax in config_shape_}
    ray = {ax: n for ax, n in total.items() if ax not in config}
    return config, ray


def _merge_axes(dims: list[int], strides: list[list[int]]):
    """Drop size-1 axes and merge adjacent axes that are contiguous in every view."""
    keep = [k for k, n in enumerate(dims) if n != 1]
    dims = [dims[k] for k in keep]
    strides = [[s[k] for k in keep] for s in strides]
    k = len(dims) - 1
    while k > 0:
        if all(s[k - 1] == s[k] * dims[k] for s in strides):
            dims[k - 1] *= dims[k]
            del dims[k]
            for s in strides:
                s[k - 1] = s[k]
                del s[k]
        k -= 1
    return dims, strides


def trace(
    system: CompiledSystem,
    rays: RayVectorArray | DeviceRays,
    accumulate: bool = False,
    axis: str | None = None,
    surf_begin: int = 0,
    surf_count: int | None = None,
    surf_step: int = 1,
    image: DeviceImage | None = None,
    image_frame=None,
    write_rays: bool = True,
    device=None,
    stats: bool = False,
    normal: na.Cartesian3dVectorArray | None = None,
    ray_axes_order: list[str] | None = None,
    _cos_log: dict | None = None,
    _exact: bool = False,
):
    """
    Trace `rays` through `system` on the device (see :func:`_trace`).  Systems with
    multilayer-coated surfaces (``MultilayerMirror`` / ``MultilayerFilm``, whose efficiency is
    ``multilayer_efficiency`` evaluated for every ray, ``optika/materials/_multilayers.py:839-935``)
    are traced in segments: up to and including a coated surface with the cosine of
    incidence captured, the stack evaluated per ray by ``optk_multilayer`` and multiplied into
    the intensity, then onwards.  Reverse traces (the stop solver) skip the coatings: they
    only use positions and directions.

    With ``system.coating == "table"`` the coatings are tabulated instead (:mod:`optika_b200._coatings`):
    efficiency(wavelength, cosine of incidence) to ``system.coating_tolerance``, looked up inside ONE
    fused launch; rays that do not reach the coating in vacuum, or a table that cannot meet the
    tolerance, take the exact chain.
    """
    if surf_count is None:
        surf_count = system.n_surface if surf_step > 0 else surf_begin + 1
    coated = [k for k in system.coatings if surf_begin <= k < surf_begin + surf_count] if surf_step > 0 else []
    if coated and not _exact and normal is None and getattr(system, "coating", "exact") == "table":
        tabled = _tabled_for_rays(system, rays, coated, surf_begin, surf_count, require_cuda(device), ray_axes_order)
        if tabled is not None:
            out = _trace(
                tabled.compiled, rays, accumulate, axis, surf_begin, surf_count, surf_step, image, image_frame,
                write_rays, device, stats, normal, ray_axes_order,
            )
            tabled.check()
            return out
    if not coated:
        return _trace(
            system, rays, accumulate, axis, surf_begin, surf_count, surf_step, image, image_frame, write_rays,
            device, stats, normal, ray_axes_order,
        )
    if normal is not None:
        raise NotImplementedError("caller-supplied normals are not supported together with coated surfaces")
    device = require_cuda(device)
    torch = _torch()
    end = surf_begin + surf_count
    states, totals = [], dict(n_rays=0, n_unvignetted=0, n_newton_iterations=0, n_binned=0)
    current, begin = rays, surf_begin
    for k in sorted(coated) + [None]:
        last = k is None
        stop = end if last else k + 1
        if stop == begin:
            break
        out = _trace(
            system, current, accumulate, axis, begin, stop - begin, 1,
            image if last else None, image_frame if last else None,
            write_rays if last else True, device, stats, None,
            ray_axes_order if current is rays else None, capture_cos=not last,
        )
        if stats:
            out, st = out
            totals.update(n_rays=st["n_rays"], n_unvignetted=st["n_unvignetted"], n_binned=st["n_binned"])
            totals["n_newton_iterations"] += st["n_newton_iterations"]
        if last:
            if out is not None:
                states.append(out)
            break
        if _cos_log is not None:  # the range of cosines this surface sees (pilot of the efficiency tables)
            finite = out.cos_incidence[torch.isfinite(out.cos_incidence)]
            if finite.numel():
                lo, hi = float(finite.min().item()), float(finite.max().item())
                old = _cos_log.get(k)
                _cos_log[k] = (lo, hi) if old is None else (min(lo, old[0]), max(hi, old[1]))
        apply_coating(system, k, out, device)
        states.append(out)
        current = out.last_state() if accumulate else out
        begin = stop
        if begin == end:
            if image is not None:  # the coated surface was the last one: bin its rays
                _trace(system, current, False, None, end, 0, 1, image, image_frame, False, device)
            break
    result = None
    if write_rays and states:
        result = states[-1] if not accumulate else DeviceRays.concatenate(states, axis if axis is not None else "surface")
    return (result, totals) if stats else result


def _pilot_indices(n: int, count: int = 5) -> np.ndarray:
    return np.unique(np.round(np.linspace(0, n - 1, min(n, count))).astype(np.int64))


def _subsample_rays(rays: RayVectorArray) -> RayVectorArray:
    """At most five samples (both ends included) along every axis of a host ray grid."""
    picks = {ax: _pilot_indices(n) for ax, n in rays.shape.items()}

    def take(a):
        if isinstance(a, (na.Cartesian2dVectorArray, na.Cartesian3dVectorArray)):
            return a._map(take)
        if not isinstance(a, na.ScalarArray):
            return a
        nd = a.ndarray
        for k, ax in enumerate(a.axes):
            if nd.shape[k] > 1:
                nd = np.take(nd, picks[ax], axis=k)
        return na.ScalarArray(nd, a.axes)

    return dataclasses.replace(
        rays, wavelength=take(rays.wavelength), position=take(rays.position), direction=take(rays.direction),
        intensity=take(rays.intensity), attenuation=take(rays.attenuation), index_refraction=take(rays.index_refraction),
        unvignetted=take(rays.unvignetted),
    )


def coating_cos_ranges(log: dict, coated, override=None) -> dict:
    """Tabulated cosine range per coated surface: what the pilot rays saw, widened by a quarter of the span."""
    ranges = {}
    for k in coated:
        if override is not None:
            ranges[k] = tuple(override[k] if isinstance(override, dict) else override)
            continue
        if k not in log:
            return None
        lo, hi = log[k]
        margin = 0.25 * (hi - lo) + 2e-3
        ranges[k] = (max(lo - margin, 1e-4), hi + margin)
    return ranges


def tabled_system_for(system: CompiledSystem, wavelengths, continuous: bool, ranges: dict, device):
    """The cached :class:`~optika_b200._coatings.TabledSystem` that covers `wavelengths` and `ranges`, or a new one."""
    from . import _coatings

    wavelengths = np.unique(np.asarray(wavelengths, dtype=float))
    tolerance = float(getattr(system, "coating_tolerance", 1e-6))
    key = (wavelengths.tobytes(), bool(continuous), str(device), tolerance)
    cache = system.__dict__.setdefault("_tabled", {})
    hit = cache.get(key)
    if hit is not None:
        if hit == "exact":
            return None
        covered = all(
            hit.tables[(k, 0)].cos_range[0] <= lo and hi <= hit.tables[(k, 0)].cos_range[1] for k, (lo, hi) in ranges.items()
        )
        if covered:
            return hit
        ranges = {k: (min(lo, hit.tables[(k, 0)].cos_range[0]), max(hi, hit.tables[(k, 0)].cos_range[1]))
                  for k, (lo, hi) in ranges.items()}
    with device_guard(device):
        tabled = _coatings.tabled_system(
            system, wavelengths, continuous, ranges, device, tolerance, int(getattr(system, "coating_max_bytes", 1 << 28))
        )
    cache[key] = tabled if tabled is not None else "exact"
    return tabled


def _vacuum_only(system: CompiledSystem, surf_begin: int, surf_count: int) -> bool:
    """No surface of the range changes the index of refraction: rays that start in vacuum stay there."""
    kinds = {system.table[c * system.n_surface + s].material_kind
             for c in range(system.n_config) for s in range(surf_begin, surf_begin + surf_count)}
    return kinds <= {L.MAT_VACUUM, L.MAT_MIRROR, L.MAT_PASS}


def _tabled_for_rays(system, rays, coated, surf_begin, surf_count, device, ray_axes_order):
    """Efficiency tables for this call, or ``None`` when the exact chain has to run."""
    from . import _coatings

    torch = _torch()
    if not _vacuum_only(system, surf_begin, surf_count):
        return None
    if isinstance(rays, DeviceRays):
        if system.n_config > 1:
            return None
        f = rays.fields
        if bool((f["index_refraction"] != 1).any().item()) or bool((f["attenuation"] != 0).any().item()):
            return None
        n = rays.size
        pick = torch.from_numpy(_pilot_indices(n, 4096)).to(f["wavelength"].device)
        pilot = DeviceRays({name: t.reshape(-1)[pick].contiguous() for name, t in f.items()},
                           rays.unvignetted.reshape(-1)[pick].contiguous(), {"_pilot": int(pick.numel())})
        finite = f["wavelength"][torch.isfinite(f["wavelength"])]
        if not finite.numel():
            return None
        wavelengths, continuous = [float(finite.min().item()), float(finite.max().item())], True
        if wavelengths[0] == wavelengths[1]:
            continuous = False
        order = None
    else:
        if np.any(np.asarray(na.as_named_array(rays.index_refraction).ndarray) != 1):
            return None
        if np.any(np.asarray(na.as_named_array(rays.attenuation).ndarray) != 0):
            return None
        w = np.unique(np.asarray(na.as_named_array(u.length(rays.wavelength)).ndarray, dtype=float))
        if not np.all(np.isfinite(w)):
            return None
        continuous = len(w) > _coatings.MAX_DISCRETE
        wavelengths = [w[0], w[-1]] if continuous else w
        pilot = _subsample_rays(rays)
        order = ray_axes_order
    # the pilot depends on the rays only: remembered per ray grid (a digest of its small named arrays)
    key = None
    if not isinstance(rays, DeviceRays):
        key = (_lowering.fingerprint(rays), surf_begin, surf_count, str(device))
    cache = system.__dict__.setdefault("_pilot_ranges", {})
    log = cache.get(key) if key is not None else None
    if log is None:
        log = {}
        trace(system, pilot, False, None, surf_begin, surf_count, 1, None, None, True, device, False, None, order,
              _cos_log=log, _exact=True)
        if key is not None:
            if len(cache) >= 16:
                cache.pop(next(iter(cache)))
            cache[key] = log
    ranges = coating_cos_ranges(log, coated, getattr(system, "coating_cos_range", None))
    if ranges is None:
        return None
    return tabled_system_for(system, wavelengths, continuous, ranges, device)


def apply_coating(system: CompiledSystem, k: int, out: "DeviceRays", device, config: int | None = None) -> None:
    """
    Multiply the per-ray multilayer efficiency of surface `k` into the (last) state of `out`.
    `out` holds every configuration of the system, or only configuration `config` when given.
    """
    material = system.coatings[k]
    n_ray = out.cos_incidence.shape[-1]
    config_dims = tuple(system.shape.values())
    indices = list(np.ndindex(*config_dims)) if config_dims else [()]
    n_config = system.n_config if config is None else 1
    todo = list(enumerate(indices)) if config is None else [(0, indices[config])]
    for c, cindex in todo:
        def last(t):
            return t.reshape(n_config, -1, n_ray)[c, -1]

        intensity = last(out.fields["intensity"])
        efficiency = material.efficiency_device(
            wavelength=last(out.fields["wavelength"]),
            cos_incidence=out.cos_incidence.reshape(n_config, n_ray)[c],
            index_refraction=last(out.fields["index_refraction"]),
            attenuation=last(out.fields["attenuation"]),
            config_shape=system.shape, cindex=tuple(cindex), device=device,
        )
        L.check(
            L.lib().optk_apply_efficiency(
                n_ray, intensity.data_ptr(), efficiency[0].data_ptr(), efficiency[1].data_ptr(), _stream_ptr(device)
            )
        )


def _trace(
    system: CompiledSystem,
    rays: RayVectorArray | DeviceRays,
    accumulate: bool = False,
    axis: str | None = None,
    surf_begin: int = 0,
    surf_count: int | None = None,
    surf_step: int = 1,
    image: DeviceImage | None = None,
    image_frame=None,
    write_rays: bool = True,
    device=None,
    stats: bool = False,
    normal: na.Cartesian3dVectorArray | None = None,
    ray_axes_order: list[str] | None = None,
    capture_cos: bool = False,
):
    """
    Trace `rays` through `system` on the device.  Returns :class:`DeviceRays`
    (or ``None`` with ``write_rays=False``), plus a stats dict when requested.

    With ``accumulate`` the result gains the axis `axis` of length `surf_count`
    right after the configuration axes (``optika/propagators.py:44-73`` stacks
    on a new leading axis; named axes make the position immaterial).
    """
    torch = _torch()
    device = require_cuda(device)
    lib = L.lib()
    if surf_count is None:
        surf_count = system.n_surface if surf_step > 0 else surf_begin + 1
    if surf_count > L.MAX_SURFACES:
        raise ValueError(f"at most {L.MAX_SURFACES} surfaces per launch; chain the trace")

    config, ray = _grid_shape(rays, system.shape, normal)
    if ray_axes_order is not None and not isinstance(rays, DeviceRays):
        # named axes have no intrinsic order: put the requested axes last, in the given
        # order (innermost = fastest varying = adjacent threads), e.g. pupil axes innermost
        # so that the rays of one warp land on the same detector pixel
        first = [ax for ax in ray if ax not in ray_axes_order]
        ray = {ax: ray[ax] for ax in first + [ax for ax in ray_axes_order if ax in ray]}
    config_axes, ray_axes = list(config), list(ray)
    n_config = int(np.prod(list(config.values()), dtype=np.int64)) if config else 1
    n_ray = int(np.prod(list(ray.values()), dtype=np.int64)) if ray else 1
    if n_config != system.n_config:
        # rays carry configuration axes the system does not vary over: they are ray axes
        raise ValueError("internal error: configuration shape mismatch")

    # ---- inputs as strided views
    if isinstance(rays, DeviceRays):
        full_axes = list(rays.shape)
        full_dims = list(rays.shape.values())
        dense = {}
        st = 1
        for ax, n in zip(reversed(full_axes), reversed(full_dims)):
            dense[ax] = 0 if n == 1 else st
            st *= n
        views = [_View(rays.fields[name], dense) for name, _ in _FIELD_GETTERS]
        mask_view = _View(rays.unvignetted, dense)
    else:
        views = [_view_of(get(rays), device) for _, get in _FIELD_GETTERS]
        unv = rays.unvignetted
        if isinstance(unv, na.ScalarArray) or not bool(unv):
            mask_view = _view_of(unv, device, is_mask=True)
        else:
            mask_view = None

    normal_views = []
    if normal is not None:
        normal_views = [_view_of(c, device) for c in (normal.x, normal.y, normal.z)]
    all_views = views + normal_views + ([mask_view] if mask_view is not None else [])
    dims = [ray[ax] for ax in ray_axes]
    strides = [[v.strides.get(ax, 0) for ax in ray_axes] for v in all_views]
    dims_m, strides_m = _merge_axes(dims, strides)
    if len(dims_m) > L.MAX_AXES:
        raise ValueError(f"ray grids with more than {L.MAX_AXES} non-mergeable axes are not supported")

    # ---- outputs
    n_states = surf_count if accumulate else 1
    out_fields = None
    out_mask = None
    if write_rays:
        out_fields = {
            name: torch.empty((n_config, n_states, n_ray), dtype=torch.float64, device=device)
            for name, _ in _FIELD_GETTERS
        }
        out_mask = torch.empty((n_config, n_states, n_ray), dtype=torch.uint8, device=device)
    out_cos = torch.empty((n_config, n_ray), dtype=torch.float64, device=device) if capture_cos else None

    stats_dev = torch.zeros(4, dtype=torch.int64, device=device) if stats else None
    frame = None
    if image_frame is not None:
        frame_shape = system.shape

    # ---- launches: one per configuration (the surface table is a kernel parameter),
    #      split along the outermost ray axis when a configuration exceeds 2^31 - 1 rays
    stream = _stream_ptr(device)
    rin = L.RaysIn()
    rout = L.RaysOut()
    config_dims = list(config.values())
    for c, cindex in enumerate(np.ndindex(*config_dims) if config_dims else [()]):
        if n_ray == 0:
            break
        base_off = [
            sum(i * v.strides.get(ax, 0) for i, ax in zip(cindex, config_axes)) for v in all_views
        ]
        if image_frame is not None:
            frame = _lowering.affine_struct(image_frame, frame_shape, tuple(cindex))
        inner = n_ray // dims_m[0] if dims_m else 1
        n0 = dims_m[0] if dims_m else 1
        step0 = max(1, min(n0, _MAX_LAUNCH // max(inner, 1)))
        if inner > _MAX_LAUNCH:
            raise ValueError("a single index of the outermost ray axis exceeds 2^31 - 1 rays")
        for i0 in range(0, n0, step0):
            m0 = min(step0, n0 - i0)
            rin.n_axes = len(dims_m)
            for a, n in enumerate(dims_m):
                rin.dims[a] = m0 if a == 0 else n
            for f, v in enumerate(views):
                off = base_off[f] + (i0 * strides_m[f][0] if dims_m else 0)
                rin.field[f] = v.tensor.data_ptr() + 8 * off
                for a in range(len(dims_m)):
                    rin.stride[f][a] = strides_m[f][a]
            for k, v in enumerate(normal_views):
                f = len(views) + k
                off = base_off[f] + (i0 * strides_m[f][0] if dims_m else 0)
                rin.normal[k] = v.tensor.data_ptr() + 8 * off
                for a in range(len(dims_m)):
                    rin.normal_stride[k][a] = strides_m[f][a]
            if mask_view is not None:
                off = base_off[-1] + (i0 * strides_m[-1][0] if dims_m else 0)
                rin.unvignetted = mask_view.tensor.data_ptr() + off
                for a in range(len(dims_m)):
                    rin.mask_stride[a] = strides_m[-1][a]
            else:
                rin.unvignetted = None
            if write_rays:
                o = c * n_states * n_ray + i0 * inner
                for f, (name, _) in enumerate(_FIELD_GETTERS):
                    rout.field[f] = out_fields[name].data_ptr() + 8 * o
                rout.unvignetted = out_mask.data_ptr() + o
            if capture_cos:
                rout.cos_incidence = out_cos.data_ptr() + 8 * (c * n_ray + i0 * inner)
            if isinstance(image, DeviceGroups):
                im = image.struct(c, first_ray=i0 * inner)  # launches start at multiples of `inner` rays
            else:
                im = image.struct(c) if image is not None else None
            L.check(
                lib.optk_trace(
                    system.handle_for(device), c, C.byref(rin), C.byref(rout) if write_rays else None,
                    surf_begin, surf_count, surf_step, 1 if accumulate else 0, n_ray,
                    C.byref(im) if im is not None else None,
                    C.byref(frame) if frame is not None else None,
                    stats_dev.data_ptr() if stats_dev is not None else None,
                    stream,
                )
            )

    result = None
    if write_rays:
        shape_ = dict(config)
        if accumulate:
            shape_[axis if axis is not None else "surface"] = n_states
        shape_.update(ray)
        result = DeviceRays(out_fields, out_mask, shape_)
        result.cos_incidence = out_cos
        result._surface_axis = len(config)
    if stats:
        s = stats_dev.cpu().numpy()
        return result, dict(
            n_rays=int(s[0]), n_unvignetted=int(s[1]), n_newton_iterations=int(s[2]), n_binned=int(s[3])
        )
    return result


# ---------------------------------------------------------------------------
# reductions over the pupil of traced rays (optk_reduce_groups)
# ---------------------------------------------------------------------------
def reduce_groups(rays: DeviceRays, axes: Sequence[str], device=None) -> dict:
    """
    Sums over the named `axes` of device-resident rays, for every index of the other axes, without
    the rays leaving the GPU (``optk_reduce_groups``; the reductions ``SequentialSystem.distortion``
    / ``vignetting`` / ``area_effective`` take over the pupil, ``optika/systems/_sequential.py:
    1266-1285, 1351-1368, 1501-1506``).  `axes` must be the innermost axes of ``rays.shape`` (the
    engine keeps the pupil axes innermost).  Returns host named arrays over the remaining axes:
    ``count`` (unvignetted rays), ``sum_intensity``, ``sum_x``, ``sum_y`` (over the unvignetted
    rays) and ``sum_x_all``, ``sum_y_all`` (over all rays).
    """
    torch = _torch()
    device = require_cuda(device)
    names = list(rays.shape)
    axes = [ax for ax in names if ax in set(axes)]
    if not axes or names[len(names) - len(axes):] != axes:
        raise ValueError(f"the reduced axes {axes} must be the innermost axes of the rays {names}")
    outer = {ax: rays.shape[ax] for ax in names[: len(names) - len(axes)]}
    n_inner = int(np.prod([rays.shape[ax] for ax in axes], dtype=np.int64))
    n_groups = int(np.prod(list(outer.values()), dtype=np.int64)) if outer else 1
    out = {
        k: torch.zeros(n_groups, dtype=torch.float64, device=device)
        for k in ("sum_intensity", "sum_x", "sum_y", "sum_x_all", "sum_y_all")
    }
    count = torch.zeros(n_groups, dtype=torch.int64, device=device)
    f = rays.fields
    # one launch holds 2^31 - 1 rays: whole groups per launch
    step = max(1, (2**31 - 1) // max(n_inner, 1))
    if n_inner > 2**31 - 1:
        raise ValueError("more than 2^31 - 1 rays per group")
    for g0 in range(0, n_groups, step):
        m = min(step, n_groups - g0)
        o = 8 * g0 * n_inner
        L.check(
            L.lib().optk_reduce_groups(
                m, n_inner, f["px"].data_ptr() + o, f["py"].data_ptr() + o, f["intensity"].data_ptr() + o,
                rays.unvignetted.data_ptr() + g0 * n_inner,
                out["sum_intensity"].data_ptr() + 8 * g0, out["sum_x"].data_ptr() + 8 * g0, out["sum_y"].data_ptr() + 8 * g0,
                count.data_ptr() + 8 * g0, out["sum_x_all"].data_ptr() + 8 * g0, out["sum_y_all"].data_ptr() + 8 * g0,
                _stream_ptr(device),
            )
        )
    dims, ax_out = tuple(outer.values()), tuple(outer)
    result = {k: na.ScalarArray(v.cpu().numpy().reshape(dims), ax_out) for k, v in out.items()}
    result["count"] = na.ScalarArray(count.cpu().numpy().reshape(dims), ax_out)
    result["n_inner"] = n_inner
    return result


# ---------------------------------------------------------------------------
# stop solver on the device (optk_solve_stops)
# ---------------------------------------------------------------------------
def solve_stops(
    surfaces,
    rays: RayVectorArray,
    variable: str,
    x0,
    y0,
    target_xy: na.Cartesian2dVectorArray,
    target: str,
    step: float,
    max_abs_error: float,
    max_iterations: int = 100,
    device=None,
):
    """
    The Newton iteration of ``SequentialSystem._calc_rayfunction_stops_only``
    (``optika/systems/_sequential.py:551-606``) in one launch per configuration.

    `surfaces` is the sub-system between the two stops (both included); `rays` start on
    ``surfaces[0]`` in global coordinates; `variable` ("direction" / "position") names the vector
    whose global x, y components are solved for, starting from `x0`, `y0`; `target` names the
    vector of the traced ray that must equal `target_xy` in the local frame of ``surfaces[-1]``.
    Returns the three components of the solved vector as named arrays, or ``None`` when the
    problem is outside what the kernel covers (the caller then iterates on the host).
    """
    if len(surfaces) < 2 or len(surfaces) > L.MAX_SURFACES:
        return None
    sag = getattr(surfaces[0], "sag", None)
    if variable == "position" and sag is not None and getattr(sag, "transformation", None) is not None:
        return None
    torch = _torch()
    device = require_cuda(device)
    system = CompiledSystem(surfaces)
    fixed = rays.position if variable == "direction" else rays.direction
    named = [rays.wavelength, fixed.x, fixed.y, fixed.z, target_xy.x, target_xy.y, x0, y0]
    named = [na.as_named_array(u.length(a) if k == 0 else a) for k, a in enumerate(named)]
    total = na.broadcast_shapes(system.shape, *[a.shape for a in named])
    full = {ax: total[ax] for ax in system.shape}
    full.update({ax: n for ax, n in total.items() if ax not in full})
    dims = tuple(full.values())
    n_config = system.n_config
    n = int(np.prod(dims, dtype=np.int64)) // max(n_config, 1) if dims else 1

    def upload(a):
        nd = np.broadcast_to(na.aligned(a, full), dims).astype(np.float64)
        return torch.from_numpy(np.ascontiguousarray(nd).reshape(n_config, n)).to(device)

    w, fx, fy, fz, tx, ty, x, y = [upload(a) for a in named]
    z = torch.empty_like(x)
    unconverged = torch.zeros(1, dtype=torch.int32, device=device)
    problem = L.StopProblem(
        variable=L.STOP_DIRECTION if variable == "direction" else L.STOP_POSITION,
        target=L.STOP_DIRECTION if target == "direction" else L.STOP_POSITION,
        surf_first=0,
        surf_last=len(surfaces) - 1,
        max_iterations=max_iterations,
        reserved=0,
        step=float(step),
        max_abs_error=float(max_abs_error),
    )
    stream = _stream_ptr(device)
    lib = L.lib()
    for c in range(n_config):
        if n == 0:
            break
        L.check(
            lib.optk_solve_stops(
                system.handle_for(device), c, C.byref(problem), n,
                w[c].data_ptr(), fx[c].data_ptr(), fy[c].data_ptr(), fz[c].data_ptr(),
                tx[c].data_ptr(), ty[c].data_ptr(), x[c].data_ptr(), y[c].data_ptr(), z[c].data_ptr(),
                unconverged.data_ptr(), stream,
            )
        )
    if int(unconverged.item()) > 0:
        raise ValueError("Max iterations exceeded")
    axes = tuple(full)
    return tuple(na.ScalarArray(t.cpu().numpy().reshape(dims), axes) for t in (x, y, z))


# ---------------------------------------------------------------------------
# unit operations of the reference API, executed by one-surface traces that are
# restricted to the relevant stages of the surface operator
# ---------------------------------------------------------------------------
@dataclasses.dataclass(eq=False)
class _Bare:
    """A surface made of exactly one element; everything else is inert."""

    sag: object = None
    material: object = None
    aperture: object = None
    rulings: object = None
    transformation: object = None

    def __post_init__(self):
        from . import sags, materials

        if self.sag is None:
            self.sag = sags.NoSag()
        if self.material is None:
            self.material = materials.Vacuum()

    @property
    def shape(self):
        return na.broadcast_shapes(
            na.shape(self.sag), na.shape(self.material), na.shape(self.aperture),
            na.shape(self.rulings), na.shape(self.transformation),
        )


@dataclasses.dataclass(eq=False)
class _FixedIndex:
    """Material of the ``snells_law`` unit operation: an explicitly given new index."""

    index: object = 1.0
    mirror: bool = False

    @property
    def is_mirror(self) -> bool:
        return self.mirror

    @property
    def shape(self):
        return na.shape(self.index)

    @property
    def transformation(self):
        return None


def _position_rays(position: na.Cartesian3dVectorArray, direction=None) -> RayVectorArray:
    if direction is None:
        direction = na.Cartesian3dVectorArray(0.0, 0.0, 1.0)
    return RayVectorArray(
        position=na.Cartesian3dVectorArray(
            u.length(position.x), u.length(position.y), u.length(position.z)
        ),
        direction=direction,
    )


def sag_intercept(sag, rays, attenuate: bool):
    """``sag.intercept`` / ``sag.propagate_rays`` (``optika/sags/_abc.py:76-122``)."""
    stages = L.STAGE_INTERCEPT | (L.STAGE_ATTENUATE if attenuate else 0)
    on_device = isinstance(rays, DeviceRays)
    out = trace(CompiledSystem([_Bare(sag=sag)], stages=stages), rays)
    return out if on_device else out.to_host()


def sag_normal(sag, position) -> na.Cartesian3dVectorArray:
    """``sag.normal(position)`` (``optika/sags/_abc.py:63-74``)."""
    out = trace(CompiledSystem([_Bare(sag=sag)], stages=L.STAGE_NORMAL_OUT), _position_rays(position))
    return out.to_host().direction


def sag_value(sag, position) -> na.ScalarArray:
    """``sag(position)`` (``optika/sags/_abc.py:48-61``)."""
    out = trace(CompiledSystem([_Bare(sag=sag)], stages=L.STAGE_SAG_OUT), _position_rays(position))
    return out.to_host().position.z


def aperture_mask(aperture, position) -> na.ScalarArray:
    """``aperture(position)`` (``optika/apertures/_apertures.py:69-80``): always tests the given vector."""
    import copy

    ap = copy.copy(aperture)
    ap.angular = False
    out = trace(CompiledSystem([_Bare(aperture=ap)], stages=L.STAGE_CLIP), _position_rays(position))
    return out.to_host().unvignetted


def aperture_clip(aperture, rays):
    """``aperture.clip_rays(rays)`` (``optika/apertures/_apertures.py:82-102``)."""
    on_device = isinstance(rays, DeviceRays)
    out = trace(CompiledSystem([_Bare(aperture=aperture)], stages=L.STAGE_CLIP), rays)
    return out if on_device else out.to_host()


def ruling_vector(spacing, position, normal) -> na.Cartesian3dVectorArray:
    """``spacing(position, normal)`` (``optika/rulings/_spacing.py:27-41``)."""
    from . import rulings as _rulings

    r = _rulings.Rulings(spacing=spacing, diffraction_order=1)
    out = trace(
        CompiledSystem([_Bare(rulings=r)], stages=L.STAGE_KAPPA_OUT), _position_rays(position), normal=normal
    )
    return out.to_host().direction


def rulings_incident_effective(rulings, rays, normal):
    """``rulings.incident_effective(rays, normal)`` (``optika/rulings/_rulings.py:170-204``)."""
    on_device = isinstance(rays, DeviceRays)
    out = trace(CompiledSystem([_Bare(rulings=rulings)], stages=L.STAGE_RULINGS), rays, normal=normal)
    return out if on_device else out.to_host()


def surface_efficiency(rays, normal, material=None, rulings=None) -> na.ScalarArray:
    """
    ``material.efficiency(rays, normal)`` / ``rulings.efficiency(rays, normal)``
    (``optika/materials/_materials.py:279-305``, ``optika/rulings/_rulings.py:205-221``): a one-surface
    trace that writes the efficiency into the intensity field (``OPTK_STAGE_EFFICIENCY_OUT``).
    """
    on_device = isinstance(rays, DeviceRays)
    system = CompiledSystem([_Bare(material=material, rulings=rulings)], stages=L.STAGE_EFFICIENCY_OUT)
    out = trace(system, rays, normal=normal)
    return out.fields["intensity"] if on_device else out.to_host().intensity


def snells_law(direction, index_refraction, index_refraction_new, normal=None, is_mirror=False):
    """``optika.materials.snells_law`` (``optika/materials/_snells_law.py:41-291``)."""
    from . import materials

    if normal is None:
        normal = na.Cartesian3dVectorArray(0.0, 0.0, -1.0)  # _snells_law.py:268-269
    material = _FixedIndex(index=index_refraction_new, mirror=bool(is_mirror))
    rays = RayVectorArray(wavelength=1.0, direction=direction, index_refraction=index_refraction)
    out = trace(CompiledSystem([_Bare(material=material)], stages=L.STAGE_REFRACT), rays, normal=normal)
    return out.to_host().direction
