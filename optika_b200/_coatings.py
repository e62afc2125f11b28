"""
Multilayer coatings inside the fused trace: efficiency tables.

``MultilayerMirror.efficiency`` / ``MultilayerFilm.efficiency`` evaluate ``multilayer_efficiency`` for
every ray (``optika/materials/_multilayers.py:839-866, 908-935``): 2.4e4 flop per ray for a 60-layer
stack, 13 x the cost of the whole uncoated trace, and a separate kernel that forces the trace to be
chained through HBM (the exact route, ``_engine.trace``).  For rays that reach the coating in vacuum the
efficiency depends on two numbers only, the wavelength and the cosine of incidence, so it can be
tabulated ONCE with the same multilayer kernel (``optk_multilayer``) and looked up inside the trace
kernel (``OPTK_EFF_TABLE2D``): one fused launch, no ray in HBM.

Accuracy is part of the contract.  Wavelength nodes are explicit:

* a ray grid with a modest number of distinct wavelengths (``raytrace`` / ``rayfunction`` /
  ``image_rays`` / ``pupil_moments`` ...) gets ONE NODE PER WAVELENGTH: the table is exact in
  wavelength, interpolated (cubic) in the cosine only;
* continuous wavelengths (``image``: stratified samples of wavelength cells, dense device rays) get a
  node on every kink of the optical constants (the ``.nk`` tables are interpolated linearly, so the
  efficiency is only piecewise smooth) and uniform refinement in between, linear interpolation.

Both axes are refined until the interpolation error, MEASURED against the exact kernel at the
midpoints of the table cells, is below ``tolerance`` (relative to the largest efficiency in the table;
default 1e-6) -- or the table would exceed ``max_bytes``, in which case there is no table and the
exact chain runs.  Rays that fall outside the tabulated cosine range are counted on the device and
make the call fail loudly instead of returning a clamped value.
"""

from __future__ import annotations
import ctypes as C
import warnings
import numpy as np
from . import named as na
from . import units as u
from . import _lib as L

__all__ = ["CoatingTable", "build_table", "tabled_system", "wavelength_nodes"]

MAX_DISCRETE = 2048  # distinct wavelengths up to which a ray grid gets one exact node per wavelength


def _torch():
    import torch

    return torch


def _chemicals_of(material) -> list:
    from .materials._multilayers import flatten_layers

    flat, _ = flatten_layers(material.layers)
    layers = list(flat)
    if material._substrate is not None:
        layers.append(material._substrate)
    return [layer._chemical for layer in layers if getattr(layer, "chemical", None) is not None]


def wavelength_nodes(material, w_lo: float, w_hi: float, per_interval: int = 1) -> np.ndarray:
    """
    Nodes for a continuous wavelength range: every tabulated wavelength of the optical constants of the
    stack's chemicals inside ``[w_lo, w_hi]`` (the efficiency has a kink there), the two ends, and
    `per_interval` - 1 equally spaced nodes inside every interval between them.
    """
    from . import chemicals

    kinks = [np.array([w_lo, w_hi])]
    for chemical in _chemicals_of(material):
        wp, _ = chemicals._load_table(chemical.file_nk)
        wp = np.asarray(wp, dtype=float)
        kinks.append(wp[(wp > w_lo) & (wp < w_hi)])
    base = np.unique(np.concatenate(kinks))
    if per_interval <= 1:
        return base
    t = np.arange(per_interval) / per_interval
    inner = (base[:-1, None] + t[None, :] * np.diff(base)[:, None]).reshape(-1)
    return np.unique(np.concatenate([inner, base[-1:]]))


class CoatingTable:
    """One table ``[n_w][n_c]`` of a coated surface at one configuration, resident on the device."""

    def __init__(self, nodes, c_first: float, c_step: float, values, error_cos: float, error_wavelength: float,
                 exact_in_wavelength: bool):
        self.nodes = nodes            # device, [n_w], ascending
        self.c_first = float(c_first)  # cosine of node 0 (one step below the tabulated range)
        self.c_step = float(c_step)
        self.values = values          # device, [n_w, n_c]
        self.error_cos = error_cos
        self.error_wavelength = error_wavelength
        self.exact_in_wavelength = exact_in_wavelength

    @property
    def n_w(self) -> int:
        return int(self.values.shape[0])

    @property
    def n_c(self) -> int:
        return int(self.values.shape[1])

    @property
    def nbytes(self) -> int:
        return 8 * self.n_w * self.n_c

    @property
    def cos_range(self) -> tuple[float, float]:
        return self.c_first + self.c_step, self.c_first + (self.n_c - 2) * self.c_step

    def lookup(self, wavelength, cosine):
        """The kernel's interpolation (``table2d_lookup`` in ``csrc/trace_impl.cuh``) in torch, for verification."""
        torch = _torch()
        nodes, table = self.nodes, self.values
        n_w, n_c = table.shape
        w = torch.clamp(wavelength, nodes[0], nodes[-1])
        i = torch.clamp(torch.searchsorted(nodes, w, right=True) - 1, 0, n_w - 2)
        lo, hi = nodes[i], nodes[i + 1]
        tw = (w - lo) / (hi - lo)
        uu = torch.clamp((cosine - self.c_first) / self.c_step, 1.0, float(n_c - 2))
        k = torch.clamp(uu.floor().long(), max=n_c - 3)
        t = uu - k
        tm, tp, t2 = t - 1.0, t + 1.0, t - 2.0
        weights = (-t * tm * t2 / 6.0, tp * tm * t2 * 0.5, -tp * t * t2 * 0.5, tp * t * tm / 6.0)

        def row(index):
            return sum(wgt * table[index, k - 1 + j] for j, wgt in enumerate(weights))

        a, b = row(i), row(i + 1)
        return a + tw * (b - a)


def _evaluate(material, wavelength, cosine, config_shape, cindex, device):
    """(s + p) / 2 of the coating for dense device tensors of wavelength and cosine (vacuum ambient)."""
    torch = _torch()
    ones = torch.ones_like(wavelength)
    out = material.efficiency_device(
        wavelength=wavelength, cos_incidence=cosine, index_refraction=ones, attenuation=torch.zeros_like(wavelength),
        config_shape=config_shape, cindex=cindex, device=device,
    )
    return 0.5 * (out[0] + out[1])


def _grid_values(material, nodes, cosines, config_shape, cindex, device):
    torch = _torch()
    w = nodes[:, None].expand(len(nodes), len(cosines)).contiguous().reshape(-1)
    c = cosines[None, :].expand(len(nodes), len(cosines)).contiguous().reshape(-1)
    return _evaluate(material, w, c, config_shape, cindex, device).reshape(len(nodes), len(cosines))


def build_table(material, wavelengths, cos_range, config_shape, cindex, device, tolerance: float = 1e-6,
                max_bytes: int = 1 << 28, continuous: bool = False):
    """
    Tabulate the coating over `wavelengths` x `cos_range`.

    `wavelengths`: the distinct wavelengths of the rays (``continuous=False``: one exact node each) or the
    two ends of a continuous range.  Returns a :class:`CoatingTable`, or ``None`` when `tolerance` cannot be
    met within `max_bytes`.
    """
    torch = _torch()
    wavelengths = np.unique(np.asarray(wavelengths, dtype=float))
    c_lo, c_hi = float(cos_range[0]), float(cos_range[1])
    if not (c_hi > c_lo):
        c_hi = c_lo + 1e-6
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)  # noqa: E731
    per_interval = 1
    while True:
        if continuous:
            nodes_host = wavelength_nodes(material, wavelengths[0], wavelengths[-1], per_interval)
        else:
            nodes_host = wavelengths
            if len(nodes_host) == 1:  # the kernel wants an interval: a second node a hair above, same physics
                nodes_host = np.array([nodes_host[0], nodes_host[0] * (1 + 1e-9)])
        nodes = up(nodes_host)
        n_cells = 16
        table = None
        while True:
            step = (c_hi - c_lo) / n_cells
            n_c = n_cells + 3  # one node beyond each end for the four-point stencil
            if 8 * len(nodes_host) * n_c > max_bytes:
                table = None
                break
            cosines = up(c_lo + step * (np.arange(n_c) - 1))
            values = _grid_values(material, nodes, cosines, config_shape, cindex, device)
            table = CoatingTable(nodes, c_lo - step, step, values, float("nan"), 0.0, not continuous)
            # error of the cubic in the cosine, measured at the cell midpoints of (a subset of) the node rows
            rows = torch.linspace(0, len(nodes_host) - 1, min(len(nodes_host), 257), device=device, dtype=torch.float64).round().long().unique()
            mid = up(c_lo + step * (np.arange(n_cells) + 0.5))
            exact = _grid_values(material, nodes[rows], mid, config_shape, cindex, device)
            w_rows = nodes[rows][:, None].expand_as(exact)
            approx = table.lookup(w_rows.reshape(-1), mid[None, :].expand_as(exact).reshape(-1)).reshape(exact.shape)
            scale = float(values.abs().max().item()) or 1.0
            table.error_cos = float((approx - exact).abs().max().item()) / scale
            if table.error_cos <= 0.5 * tolerance:
                break
            n_cells *= 2
        if table is None:
            return None
        if not continuous:
            return table
        # error of the linear interpolation between wavelength nodes, at the interval midpoints
        mid_w = 0.5 * (nodes[1:] + nodes[:-1])
        pick = torch.linspace(0, len(mid_w) - 1, min(len(mid_w), 4097), device=device, dtype=torch.float64).round().long().unique()
        cos_probe = up(np.linspace(c_lo, c_hi, 9))
        exact = _grid_values(material, mid_w[pick], cos_probe, config_shape, cindex, device)
        approx = table.lookup(
            mid_w[pick][:, None].expand_as(exact).reshape(-1), cos_probe[None, :].expand_as(exact).reshape(-1)
        ).reshape(exact.shape)
        scale = float(table.values.abs().max().item()) or 1.0
        table.error_wavelength = float((approx - exact).abs().max().item()) / scale
        if table.error_wavelength <= 0.5 * tolerance:
            return table
        # linear interpolation: the error falls with the square of the spacing
        grow = max(2, int(np.ceil(np.sqrt(table.error_wavelength / (0.4 * tolerance)))))
        per_interval *= grow
        if 8 * (len(wavelength_nodes(material, wavelengths[0], wavelengths[-1], 1)) * per_interval) * table.n_c > max_bytes:
            return None


class TabledSystem:
    """
    A compiled system whose coated surfaces carry efficiency tables: traced by ONE fused launch.  Holds
    the tables, the out-of-range counter and the derived :class:`~optika_b200._engine.CompiledSystem`.
    """

    def __init__(self, compiled, tables: dict, counter, device):
        self.compiled = compiled
        self.tables = tables  # {(surface, config): CoatingTable}
        self.counter = counter
        self.device = device

    def check(self) -> None:
        """Raise when a ray of the launches so far fell outside the tabulated ranges (synchronises)."""
        n = int(self.counter.item())
        if n:
            self.counter.zero_()
            ranges = {k: t.cos_range for k, t in self.tables.items()}
            raise ValueError(
                f"{n} rays reached a coated surface outside its efficiency table (cosine ranges {ranges}); "
                "pass a wider `coating_cos_range` or use the exact per-ray evaluation (coating='exact')"
            )

    @property
    def errors(self) -> dict:
        return {k: (t.error_cos, t.error_wavelength) for k, t in self.tables.items()}


def tabled_system(system, wavelengths, continuous: bool, cos_ranges: dict, device, tolerance: float = 1e-6,
                  max_bytes: int = 1 << 28):
    """
    A :class:`TabledSystem` for `system` (a ``CompiledSystem`` with coatings): every coated surface gets a
    table over `wavelengths` (distinct values, or the ends of a continuous range) and ``cos_ranges[surface]``.
    ``None`` when some table cannot meet `tolerance` within `max_bytes` (a warning says so).
    """
    from . import _engine, _lowering

    torch = _torch()
    config_dims = tuple(system.shape.values())
    indices = list(np.ndindex(*config_dims)) if config_dims else [()]
    counter = torch.zeros(1, dtype=torch.int64, device=device)
    table, shape_ = _lowering.lower_system(system.surfaces)
    n_surface = len(system.surfaces)
    tables = {}
    for k, material in system.coatings.items():
        varies = bool(na.shape(material))  # the stack has configuration axes: one table per configuration
        shared = None
        for c, cindex in enumerate(indices):
            if shared is None or varies:
                shared = build_table(
                    material, wavelengths, cos_ranges[k], system.shape, tuple(cindex), device, tolerance, max_bytes,
                    continuous,
                )
                if shared is None:
                    warnings.warn(
                        f"no efficiency table within {max_bytes} bytes meets the tolerance {tolerance:g} for the coating "
                        f"of surface {k}; evaluating it exactly per ray"
                    )
                    return None
            tables[(k, c)] = shared
            S = table[c * n_surface + k]
            S.material_efficiency = L.EFF_TABLE2D
            S.material_lut_n = shared.n_w
            S.material_lut_x = shared.nodes.data_ptr()
            S.material_lut_y = shared.values.data_ptr()
            S.material[0] = shared.c_first
            S.material[1] = 1.0 / shared.c_step
            S.material[2] = float(shared.n_c)
            S.ruling_lut_x = counter.data_ptr()
    local_last = bool(system.table[n_surface - 1].flags & L.F_LOCAL_OUT) if n_surface else False
    compiled = _engine.CompiledSystem(system.surfaces, local_last=local_last, lowered=(table, shape_))
    compiled.coatings = {}  # the efficiency is in the table now: no chaining
    return TabledSystem(compiled, tables, counter, device)
