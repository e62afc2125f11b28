"""
On-device ray grids: the host description of ``optk_grid_t`` and its launcher.

``SequentialSystem.image`` (``optika/systems/_sequential.py:1088-1206``) samples one
stratified random ray per cell of a (wavelength, field, pupil) vertex grid
(``_rayfunction_from_vertices``, ``:1002-1086``) and weights it with
``radiance * cell_area``.  For the separable grids the reference's own examples use, all of
that is a few small 1-D / 2-D arrays; :func:`trace_grid` hands them to
``optk_trace_grid``, which draws, traces and bins every ray inside one kernel launch.
"""

from __future__ import annotations
import ctypes as C
import dataclasses
import numpy as np
from . import _lib as L
from . import _engine

__all__ = ["RayGrid", "trace_grid", "AXES"]

AXES = ("wavelength", "field_x", "field_y", "pupil_x", "pupil_y")
LAUNCHES = 0  # kernels launched by trace_grid in this process (bench accounting)
_MAX_LAUNCH = 2**31 - 1


@dataclasses.dataclass(eq=False)
class RayGrid:
    """
    A separable grid of cell vertices in device units (mm, rad).

    ``vertices``: five 1-D arrays (wavelength, field_x, field_y, pupil_x, pupil_y) of
    ``n + 1`` values; ``weight_scene[n0, n1, n2]`` and ``weight_pupil[n3, n4]`` multiply
    into the intensity of every ray; ``frame = (R[3, 3], t[3])`` maps the generated rays to
    the coordinates of the first surface; ``begin`` / ``count`` select a sub-box (the
    random stream does not depend on it, see ``include/optk.h``).
    """

    vertices: tuple
    at_infinity: bool = True
    weight_scene: np.ndarray | None = None
    weight_pupil: np.ndarray | None = None
    jitter: bool = True
    seed: int = 0
    frame: tuple | None = None
    axes: tuple = AXES
    begin: tuple | None = None
    count: tuple | None = None
    chromatic: tuple = ()
    """Axes (1 .. 4) whose vertex array is 2-D ``[n_wavelength + 1][n + 1]``: one row per wavelength vertex."""

    def __post_init__(self):
        self.vertices = tuple(np.ascontiguousarray(v, dtype=np.float64) for v in self.vertices)
        if len(self.vertices) != 5:
            raise ValueError("a ray grid has five axes: wavelength, field x/y, pupil x/y")
        if self.vertices[0].ndim != 1:
            raise ValueError("the wavelength vertices must be one-dimensional")
        self.chromatic = tuple(sorted(int(a) for a in self.chromatic))
        n_w = len(self.vertices[0])
        for a in self.chromatic:
            if a not in (1, 2, 3, 4) or self.vertices[a].ndim != 2 or self.vertices[a].shape[0] != n_w:
                raise ValueError(
                    f"chromatic axis {a} needs a vertex array [n_wavelength + 1 = {n_w}][n + 1], got {self.vertices[a].shape}"
                )
        for name, (a, b) in (("field", (1, 2)), ("pupil", (3, 4))):
            va, vb = self.vertices[a], self.vertices[b]
            if a in self.chromatic or b in self.chromatic:
                if any(self.vertices[c].ndim != 1 for c in (a, b) if c not in self.chromatic):
                    raise ValueError(f"the {name} axis paired with a chromatic one must be 1-D (separable)")
                continue
            if not ((va.ndim == 1 and vb.ndim == 1) or (va.ndim == 2 and va.shape == vb.shape)):
                raise ValueError(
                    f"the {name} vertices are two 1-D arrays (separable grid) or two 2-D arrays of one shape "
                    f"(curvilinear grid), got shapes {va.shape} and {vb.shape}"
                )
        n = self.n
        if any(k < 1 for k in n):
            raise ValueError(f"every axis needs at least two vertices, got cells {n}")
        if self.weight_scene is not None:
            self.weight_scene = np.ascontiguousarray(np.broadcast_to(self.weight_scene, n[:3]), dtype=np.float64)
        if self.weight_pupil is not None:
            wp = np.asarray(self.weight_pupil, dtype=np.float64)
            target = (n[0],) + tuple(n[3:]) if wp.ndim == 3 else tuple(n[3:])  # 3-D: one set of areas per wavelength cell
            self.weight_pupil = np.ascontiguousarray(np.broadcast_to(wp, target), dtype=np.float64)
        if self.begin is None:
            self.begin = (0,) * 5
        if self.count is None:
            self.count = tuple(n[a] - self.begin[a] for a in range(5))
        self._device = {}

    @property
    def field_2d(self) -> bool:
        return self.vertices[1].ndim == 2 and not ({1, 2} & set(self.chromatic))

    @property
    def pupil_2d(self) -> bool:
        return self.vertices[3].ndim == 2 and not ({3, 4} & set(self.chromatic))

    @property
    def n(self) -> tuple:
        v = self.vertices

        def cells(a, b, curvilinear):
            if curvilinear:
                return (v[a].shape[0] - 1, v[a].shape[1] - 1)
            return tuple(v[c].shape[-1] - 1 for c in (a, b))

        return (len(v[0]) - 1,) + cells(1, 2, self.field_2d) + cells(3, 4, self.pupil_2d)

    @property
    def shape(self) -> dict[str, int]:
        return dict(zip(self.axes, self.count))

    @property
    def size(self) -> int:
        return int(np.prod(self.count, dtype=np.int64))

    def sub(self, begin, count) -> "RayGrid":
        """The same grid restricted to a sub-box (absolute cell indices)."""
        g = dataclasses.replace(self, begin=tuple(begin), count=tuple(count))
        g._device = self._device
        return g

    def shard(self, rank: int, world_size: int, axis: int = 3) -> "RayGrid":
        """Contiguous slab of this grid's box for one rank (default: along pupil_x)."""
        from .distributed import slab

        s = slab(self.count[axis], rank, world_size)
        begin, count = list(self.begin), list(self.count)
        begin[axis] += s.start
        count[axis] = s.stop - s.start
        return self.sub(begin, count)

    def on_device(self, device):
        torch = _engine._torch()
        key = str(device)
        if key not in self._device:
            up = lambda a: None if a is None else torch.from_numpy(np.array(a, dtype=np.float64).reshape(-1)).to(device)  # noqa: E731
            self._device[key] = dict(
                vertices=[up(v) for v in self.vertices],
                weight_scene=up(self.weight_scene),
                weight_pupil=up(self.weight_pupil),
                angular_cells=[up(c) for c in self.angular_cells()],
            )
        return self._device[key]

    def angular_cells(self) -> list:
        """
        ``optk_grid_t.angular_cells``: for the two angular axes (field for an object at infinity, else
        pupil) one record ``{sin v_i, cos v_i, v_{i+1} - v_i, 0}`` per cell, when the grid is separable
        along them and no cell is wider than 0.01 rad; else ``[None, None]`` (full sincos per ray).
        """
        a, b = (1, 2) if self.at_infinity else (3, 4)
        va, vb = self.vertices[a], self.vertices[b]
        if va.ndim != 1 or vb.ndim != 1 or self.chromatic:
            return [None, None]
        if max(np.max(np.abs(np.diff(va))), np.max(np.abs(np.diff(vb)))) > 0.01:
            return [None, None]
        if not (np.all(np.isfinite(va)) and np.all(np.isfinite(vb))):
            return [None, None]

        def pack(v):
            out = np.zeros((len(v) - 1, 4))
            out[:, 0], out[:, 1], out[:, 2] = np.sin(v[:-1]), np.cos(v[:-1]), np.diff(v)
            return out

        return [pack(va), pack(vb)]

    def struct(self, device, begin=None, count=None) -> L.Grid:
        dev = self.on_device(device)
        g = L.Grid()
        g.n[:] = self.n
        g.begin[:] = self.begin if begin is None else begin
        g.count[:] = self.count if count is None else count
        g.at_infinity = 1 if self.at_infinity else 0
        g.jitter = 1 if self.jitter else 0
        g.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        for a in range(5):
            g.vertices[a] = dev["vertices"][a].data_ptr()
        g.weight_scene = None if dev["weight_scene"] is None else dev["weight_scene"].data_ptr()
        g.weight_pupil = None if dev["weight_pupil"] is None else dev["weight_pupil"].data_ptr()
        g.field_2d = 1 if self.field_2d else 0
        g.pupil_2d = 1 if self.pupil_2d else 0
        g.chromatic = sum(1 << a for a in self.chromatic)
        g.weight_pupil_chromatic = 1 if (self.weight_pupil is not None and self.weight_pupil.ndim == 3) else 0
        for k in range(2):
            cells = dev["angular_cells"][k]
            g.angular_cells[k] = None if cells is None else cells.data_ptr()
        if self.frame is not None:
            g.has_frame = 1
            g.frame.r[:] = list(np.asarray(self.frame[0], dtype=float).reshape(9))
            g.frame.t[:] = list(np.asarray(self.frame[1], dtype=float).reshape(3))
        return g


def _boxes(begin, count, limit):
    """Split a box into sub-boxes of at most `limit` cells, cutting the outermost axes first."""
    begin, count = list(begin), list(count)
    total = int(np.prod(count, dtype=np.int64))
    if total <= limit:
        yield tuple(begin), tuple(count)
        return
    for a in range(5):
        if count[a] > 1:
            inner = total // count[a]
            step = max(1, limit // inner)
            for i in range(0, count[a], step):
                b, c = list(begin), list(count)
                b[a] += i
                c[a] = min(step, count[a] - i)
                yield from _boxes(b, c, limit)
            return
    raise ValueError("cannot split the grid")  # pragma: no cover


def trace_grid(
    system: "_engine.CompiledSystem",
    grid: RayGrid,
    config: int = 0,
    accumulate: bool = False,
    axis: str | None = None,
    surf_begin: int = 0,
    surf_count: int | None = None,
    surf_step: int = 1,
    image: "_engine.DeviceImage | None" = None,
    image_plane: int | None = None,
    image_frame=None,
    write_rays: bool = True,
    device=None,
    stats: bool = False,
    max_launch: int = _MAX_LAUNCH,
    capture_cos: bool = False,
):
    """
    Generate the rays of `grid` on the device and trace them through configuration
    `config` of `system` (``optk_trace_grid``).  Returns :class:`DeviceRays` over the
    grid's box (or ``None`` with ``write_rays=False``), plus a stats dict when requested.
    Coated surfaces are NOT applied here (see :func:`trace_grid_coated`).
    """
    torch = _engine._torch()
    device = _engine.require_cuda(device)
    lib = L.lib()
    if surf_count is None:
        surf_count = system.n_surface if surf_step > 0 else surf_begin + 1
    n_ray = grid.size
    n_states = surf_count if accumulate else 1
    out_fields = out_mask = None
    if write_rays:
        out_fields = {
            name: torch.empty((n_states, n_ray), dtype=torch.float64, device=device)
            for name, _ in _engine._FIELD_GETTERS
        }
        out_mask = torch.empty((n_states, n_ray), dtype=torch.uint8, device=device)
    out_cos = torch.empty(n_ray, dtype=torch.float64, device=device) if capture_cos else None
    stats_dev = torch.zeros(4, dtype=torch.int64, device=device) if stats else None
    frame = None
    if image_frame is not None:
        frame = L.Affine()
        frame.r[:] = list(np.asarray(image_frame[0], dtype=float).reshape(9))
        frame.t[:] = list(np.asarray(image_frame[1], dtype=float).reshape(3))
    im = image.struct(config if image_plane is None else image_plane) if image is not None else None
    stream = _engine._stream_ptr(device)
    rout = L.RaysOut()
    offset = 0
    for begin, count in _boxes(grid.begin, grid.count, max_launch) if n_ray else ():
        n_box = int(np.prod(count, dtype=np.int64))
        if write_rays:
            # sub-boxes are contiguous in the output only when they are slabs of the
            # outermost non-trivial axis
            lead = [a for a in range(5) if grid.count[a] > 1]
            if any(count[a] != grid.count[a] for a in lead[1:]):
                raise ValueError("ray output needs slabs of the outermost axis; lower the inner axes or skip write_rays")
            for f, (name, _) in enumerate(_engine._FIELD_GETTERS):
                rout.field[f] = out_fields[name].data_ptr() + 8 * offset
            rout.unvignetted = out_mask.data_ptr() + offset
            if capture_cos:
                rout.cos_incidence = out_cos.data_ptr() + 8 * offset
        g = grid.struct(device, begin, count)
        global LAUNCHES
        LAUNCHES += 1
        L.check(
            lib.optk_trace_grid(
                system.handle_for(device), config, C.byref(g), C.byref(rout) if write_rays else None,
                surf_begin, surf_count, surf_step, 1 if accumulate else 0, n_ray,
                C.byref(im) if im is not None else None,
                C.byref(frame) if frame is not None else None,
                stats_dev.data_ptr() if stats_dev is not None else None,
                stream,
            )
        )
        offset += n_box
    result = None
    if write_rays:
        shape_ = {}
        if accumulate:
            shape_[axis if axis is not None else "surface"] = n_states
        shape_.update(grid.shape)
        result = _engine.DeviceRays(out_fields, out_mask, shape_)
        result.cos_incidence = out_cos
    if stats:
        s = stats_dev.cpu().numpy()
        return result, dict(
            n_rays=int(s[0]), n_unvignetted=int(s[1]), n_newton_iterations=int(s[2]), n_binned=int(s[3])
        )
    return result


def _trace_dense(system, config: int, rays, surf_begin: int, surf_count: int, image, write_rays: bool,
                 capture_cos: bool, device):
    """One ``optk_trace`` launch of configuration `config` on dense device rays (a link of the coated chain)."""
    torch = _engine._torch()
    n = rays.size
    rin = L.RaysIn()
    rin.n_axes = 1
    rin.dims[0] = n
    for f, (name, _) in enumerate(_engine._FIELD_GETTERS):
        rin.field[f] = rays.fields[name].data_ptr()
        rin.stride[f][0] = 1
    rin.unvignetted = rays.unvignetted.data_ptr()
    rin.mask_stride[0] = 1
    rout = L.RaysOut()
    result = None
    if write_rays:
        fields = {name: torch.empty(n, dtype=torch.float64, device=device) for name, _ in _engine._FIELD_GETTERS}
        mask = torch.empty(n, dtype=torch.uint8, device=device)
        for f, (name, _) in enumerate(_engine._FIELD_GETTERS):
            rout.field[f] = fields[name].data_ptr()
        rout.unvignetted = mask.data_ptr()
        result = _engine.DeviceRays(fields, mask, dict(rays.shape))
        if capture_cos:
            result.cos_incidence = torch.empty(n, dtype=torch.float64, device=device)
            rout.cos_incidence = result.cos_incidence.data_ptr()
    im = image.struct(config) if image is not None else None
    global LAUNCHES
    LAUNCHES += 1
    L.check(
        L.lib().optk_trace(
            system.handle_for(device), config, C.byref(rin), C.byref(rout) if write_rays else None, surf_begin, surf_count, 1,
            0, 0, C.byref(im) if im is not None else None, None, None, _engine._stream_ptr(device),
        )
    )
    return result


def _chain(system, sub: RayGrid, config: int, image, device, cos_log: dict | None = None) -> None:
    """
    The exact route for one box of a grid: generated and traced to the first coated surface, the coating
    evaluated per ray (:func:`optika_b200._engine.apply_coating`), and so on; the last link bins into
    `image` (skipped without one).  `cos_log` collects the range of cosines every coated surface sees.
    """
    torch = _engine._torch()
    coated = sorted(system.coatings)
    n_surface = system.n_surface

    def log(k, rays):
        if cos_log is None:
            return
        finite = rays.cos_incidence[torch.isfinite(rays.cos_incidence)]
        if finite.numel():
            lo, hi = float(finite.min().item()), float(finite.max().item())
            old = cos_log.get(k)
            cos_log[k] = (lo, hi) if old is None else (min(lo, old[0]), max(hi, old[1]))

    rays = trace_grid(system, sub, config=config, surf_count=coated[0] + 1, device=device, capture_cos=True)
    log(coated[0], rays)
    _engine.apply_coating(system, coated[0], rays, device, config=config)
    at = coated[0] + 1
    for k in coated[1:] + [None]:
        stop = n_surface if k is None else k + 1
        final = k is None
        if final and image is None:
            break
        rays = _trace_dense(system, config, rays, at, stop - at, image if final else None, not final, not final, device)
        if not final:
            log(k, rays)
            _engine.apply_coating(system, k, rays, device, config=config)
        at = stop


def _tabled_for_grid(system, grid: RayGrid, device):
    """Efficiency tables that cover `grid` in every configuration, or ``None`` (exact chain)."""
    coated = sorted(system.coatings)
    if not _engine._vacuum_only(system, 0, system.n_surface):
        return None
    w = grid.vertices[0]
    if grid.jitter:
        wavelengths, continuous = [float(w.min()), float(w.max())], True
    else:
        wavelengths, continuous = 0.5 * (w[:-1] + w[1:]), False
    cache = system.__dict__.setdefault("_grid_cos_ranges", {})
    key = (grid.vertices[0].tobytes(), tuple(v.tobytes() for v in grid.vertices[1:]), grid.jitter, grid.seed)
    ranges = cache.get(key)
    if ranges is None:
        # pilot: at most five vertices (both ends included) along every axis, every configuration
        def pick(v):
            if v.ndim == 1:
                return v[_engine._pilot_indices(len(v))]
            return v[np.ix_(_engine._pilot_indices(v.shape[0]), _engine._pilot_indices(v.shape[1]))]

        pilot = RayGrid(
            vertices=tuple(pick(v) for v in grid.vertices), at_infinity=grid.at_infinity, jitter=grid.jitter,
            seed=grid.seed, frame=grid.frame, axes=grid.axes, chromatic=grid.chromatic,
        )
        log = {}
        for c in range(system.n_config):
            _chain(system, pilot, c, None, device, cos_log=log)
        ranges = _engine.coating_cos_ranges(log, coated, getattr(system, "coating_cos_range", None))
        if ranges is None:
            return None
        cache[key] = ranges
    return _engine.tabled_system_for(system, wavelengths, continuous, ranges, device)


def trace_grid_coated(system, grid: RayGrid, config: int, image, device=None, max_rays: int = 1 << 25) -> None:
    """
    Fused image of a system with multilayer-coated surfaces.

    ``system.coating == "table"``: the coatings are efficiency tables (:mod:`optika_b200._coatings`) and the
    whole grid is ONE fused launch like an uncoated system; the caller checks ``system._tabled_pending``
    (rays outside a table) once its launches are done.  Otherwise -- or when the rays do not reach the
    coating in vacuum, or no table meets the tolerance -- the exact route: the grid is cut into boxes of
    at most `max_rays` rays (they do live in HBM between the links of the chain), see :func:`_chain`.
    """
    device = _engine.require_cuda(device)
    if grid.size == 0:
        return  # an empty slab (more ranks than cells): nothing to trace, the caller still reduces
    if getattr(system, "coating", "exact") == "table":
        tabled = _tabled_for_grid(system, grid, device)
        if tabled is not None:
            trace_grid(tabled.compiled, grid, config=config, image=image, write_rays=False, device=device)
            pending = system.__dict__.setdefault("_tabled_pending", [])
            if tabled not in pending:
                pending.append(tabled)
            return
    for begin, count in _boxes(grid.begin, grid.count, max_rays):
        _chain(system, grid.sub(begin, count), config, image, device)
