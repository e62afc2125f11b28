"""
Multilayer reflectivity and transmissivity.

``multilayer_efficiency`` keeps the signature of
``optika.materials.multilayer_efficiency`` (``optika/materials/_multilayers.py:240-249``)
and returns ``(R, T)`` as :class:`~optika_b200.vectors.PolarizationVectorArray`.
The host interpolates the optical constants of each layer once per wavelength
grid (``optika/chemicals/_chemicals.py:101-144``) and describes the broadcast of
(wavelength, direction, ambient index, thicknesses, interface widths) by named
axes; the transfer-matrix chain runs in ``optika_b200/csrc/multilayer.cu``.
"""

from __future__ import annotations
import ctypes as C
import numpy as np
from .. import named as na
from .. import units as u
from .. import _lib as L
from ..vectors import PolarizationVectorArray
from ._layers import AbstractLayer, Layer, LayerSequence, PeriodicLayerSequence

__all__ = ["multilayer_efficiency", "multilayer_efficiency_device", "flatten_layers"]


def flatten_layers(layers) -> tuple[list[Layer], list[tuple[int, int, int]]]:
    """
    Flatten nested layer containers into a list of :class:`Layer` plus segments
    ``(first, count, repeat)``; a periodic sequence of plain layers becomes one
    repeated segment (``optika/materials/_layers.py:611-645``), anything more
    deeply nested is unrolled (``layer_sequence``, ``:606-609``).
    """
    flat: list[Layer] = []
    segments: list[tuple[int, int, int]] = []

    def add_run(items):
        first = len(flat)
        flat.extend(items)
        if segments and segments[-1][2] == 1 and segments[-1][0] + segments[-1][1] == first:
            segments[-1] = (segments[-1][0], segments[-1][1] + len(items), 1)
        else:
            segments.append((first, len(items), 1))

    def visit(item):
        if item is None:
            return
        if isinstance(item, Layer):
            add_run([item])
        elif isinstance(item, PeriodicLayerSequence):
            if all(isinstance(x, Layer) for x in item.layers) and item.num_periods >= 1:
                first = len(flat)
                flat.extend(item.layers)
                segments.append((first, len(item.layers), int(item.num_periods)))
            else:
                for _ in range(int(item.num_periods)):
                    for x in item.layers:
                        visit(x)
        elif isinstance(item, LayerSequence):
            for x in item.layers:
                visit(x)
        elif isinstance(item, (list, tuple)):
            for x in item:
                visit(x)
        else:
            raise TypeError(f"unsupported layer type {type(item)}")

    visit(layers)
    return flat, segments


def _strided(value, axes: list[str], device, torch):
    """Upload `value` (scalar / named array) and return (tensor, strides over `axes`)."""
    if isinstance(value, na.ScalarArray):
        nd = np.ascontiguousarray(np.asarray(value.ndarray, dtype=np.float64))
        own = {ax: (0 if n == 1 else st // 8) for ax, st, n in zip(value.axes, nd.strides, nd.shape)}
        t = torch.from_numpy(nd.reshape(-1).copy()).to(device)
        return t, [own.get(ax, 0) for ax in axes]
    t = torch.from_numpy(np.asarray(value, dtype=np.float64).reshape(1).copy()).to(device)
    return t, [0] * len(axes)


def _split_complex(value):
    if isinstance(value, na.ScalarArray):
        nd = np.asarray(value.ndarray)
        re = na.ScalarArray(np.real(nd).astype(np.float64), value.axes)
        im = na.ScalarArray(np.imag(nd).astype(np.float64), value.axes) if np.iscomplexobj(nd) else None
        return re, im
    v = complex(value)
    return v.real, (v.imag if v.imag != 0 else None)


def multilayer_efficiency_device(
    wavelength,
    direction=1,
    n=1,
    layers: None | AbstractLayer | list = None,
    substrate: None | Layer = None,
    device=None,
) -> tuple[PolarizationVectorArray, PolarizationVectorArray]:
    """
    Device-resident variant of :func:`multilayer_efficiency`: returns
    ``(tensor[4, n_eval], axes, dims)`` with rows R_s, R_p, T_s, T_p left in HBM.
    """
    from .. import _engine

    torch = _engine._torch()
    device = _engine.require_cuda(device)

    if substrate is None:
        substrate = Layer()  # vacuum, _multilayers.py:487-490
    flat, segments = flatten_layers(layers)
    if len(flat) + 1 > L.ML_MAX_LAYERS:
        # unroll-free limit of the shared-memory layer table
        raise ValueError(f"at most {L.ML_MAX_LAYERS - 1} distinct layers are supported")
    if len(segments) > 32:
        raise ValueError("at most 32 layer segments are supported")

    wavelength = na.as_named_array(u.length(wavelength)) if not isinstance(wavelength, na.ScalarArray) else wavelength
    dir_re, dir_im = _split_complex(direction)
    n_re, n_im = _split_complex(n)

    stack = flat + [substrate]
    n_layers = [layer.n(wavelength) for layer in stack]
    thickness = [0 if layer.thickness is None else u.length(layer.thickness) for layer in stack]
    widths = [None if layer.interface is None else u.length(layer.interface.width) for layer in stack]

    shape_ = na.shape_broadcasted(
        wavelength, dir_re, dir_im, n_re, n_im, *n_layers, *thickness, *[w for w in widths if w is not None]
    )
    axes = list(shape_)
    dims = [shape_[ax] for ax in axes]
    # merge / drop axes so that at most OPTK_ML_MAX_AXES remain
    keep = [k for k, d in enumerate(dims) if d != 1]
    axes = [axes[k] for k in keep]
    dims = [dims[k] for k in keep]
    if len(axes) > L.ML_MAX_AXES:
        raise ValueError(f"evaluation grids with more than {L.ML_MAX_AXES} axes are not supported")
    n_eval = int(np.prod(dims, dtype=np.int64)) if dims else 1

    keepalive = []

    def view(value):
        t, st = _strided(value, axes, device, torch)
        keepalive.append(t)
        return t, st

    inp = L.MlInput()
    inp.n_axes = len(axes)
    for a, d in enumerate(dims):
        inp.dims[a] = d
    t, st = view(wavelength)
    inp.wavelength = t.data_ptr()
    inp.wavelength_stride[:] = st + [0] * (L.ML_MAX_AXES - len(st))
    t, st = view(dir_re)
    inp.direction_re = t.data_ptr()
    inp.direction_stride[:] = st + [0] * (L.ML_MAX_AXES - len(st))
    if dir_im is not None:
        t, st2 = view(dir_im)
        inp.direction_im = t.data_ptr()
    t, st = view(n_re)
    inp.n_re = t.data_ptr()
    inp.n_stride[:] = st + [0] * (L.ML_MAX_AXES - len(st))
    if n_im is not None:
        t, st2 = view(n_im)
        inp.n_im = t.data_ptr()

    table = (L.MlLayer * len(stack))()
    for j, layer in enumerate(stack):
        re, im = _split_complex(n_layers[j])
        t, st = view(re)
        table[j].n_re = t.data_ptr()
        table[j].n_stride[:] = st + [0] * (L.ML_MAX_AXES - len(st))
        if im is not None:
            t, _ = view(im)
            table[j].n_im = t.data_ptr()
        t, st = view(thickness[j])
        table[j].thickness = t.data_ptr()
        table[j].thickness_stride[:] = st + [0] * (L.ML_MAX_AXES - len(st))
        if widths[j] is not None:
            t, st = view(widths[j])
            table[j].width = t.data_ptr()
            table[j].width_stride[:] = st + [0] * (L.ML_MAX_AXES - len(st))
            table[j].profile_kind = layer.interface.kind
    segs = (L.MlSegment * max(len(segments), 1))()
    for g, (first, count, repeat) in enumerate(segments):
        segs[g].first, segs[g].count, segs[g].repeat = first, count, repeat

    out = torch.empty((4, n_eval), dtype=torch.float64, device=device)
    ptrs = [out.data_ptr() + 8 * n_eval * k for k in range(4)]
    L.check(
        L.lib().optk_multilayer(
            C.byref(inp), len(stack), table, len(segments), segs,
            ptrs[0], ptrs[1], ptrs[2], ptrs[3], _engine._stream_ptr(device),
        )
    )
    torch.cuda.current_stream(device).synchronize()
    del keepalive
    return out, axes, dims


def multilayer_efficiency(
    wavelength,
    direction=1,
    n=1,
    layers: None | AbstractLayer | list = None,
    substrate: None | Layer = None,
    device=None,
) -> tuple[PolarizationVectorArray, PolarizationVectorArray]:
    """
    Reflectivity and transmissivity of a multilayer stack for s and p
    polarisation (``optika/materials/_multilayers.py:240-532``).

    Parameters keep the reference's meaning: `wavelength` in vacuum (mm),
    `direction` the cosine of the incidence angle in the ambient medium, `n` the
    (complex) ambient index, `layers` from the ambient side down, `substrate`
    the medium below (its thickness is ignored).
    """
    out, axes, dims = multilayer_efficiency_device(wavelength, direction, n, layers, substrate, device)
    host = out.cpu().numpy().reshape([4] + dims)

    def wrap(a):
        return na.ScalarArray(a, tuple(axes))

    reflectivity = PolarizationVectorArray(s=wrap(host[0]), p=wrap(host[1]))
    transmissivity = PolarizationVectorArray(s=wrap(host[2]), p=wrap(host[3]))
    return reflectivity, transmissivity


# ---------------------------------------------------------------------------
# per-ray evaluation: multilayer coatings as surface materials
# ---------------------------------------------------------------------------
_nk_device = {}


def _nk_table(chemical, device, torch):
    """The ``.nk`` table of a chemical on the device: (wavelength [mm], n, k), cached."""
    from .. import chemicals

    key = (chemical.file_nk, str(device))
    if key not in _nk_device:
        wp, fp = chemicals._load_table(chemical.file_nk)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)  # noqa: E731
        _nk_device[key] = (up(wp), up(fp.real), up(fp.imag), len(wp))
    return _nk_device[key]


def multilayer_efficiency_rays(
    wavelength, cos_incidence, index_refraction, attenuation, layers, substrate, config_shape, cindex, device,
):
    """
    ``multilayer_efficiency`` for N rays whose wavelength, cosine of incidence, ambient index
    and attenuation are dense device tensors (``optika/materials/_multilayers.py:852-865,
    921-934``): ``n = index_refraction + i attenuation wavelength / (4 pi)``, the optical
    constants of every layer interpolated at the ray wavelengths on the device
    (``optk_interp``), thicknesses / interface widths taken at configuration `cindex`.
    Returns a device tensor ``[4, N]``: R_s, R_p, T_s, T_p.
    """
    from .. import _engine, _lowering

    torch = _engine._torch()
    lib = L.lib()
    stream = _engine._stream_ptr(device)
    n_ray = int(wavelength.numel())
    if substrate is None:
        substrate = Layer()
    flat, segments = flatten_layers(layers)
    stack = flat + [substrate]
    if len(stack) > L.ML_MAX_LAYERS:
        raise ValueError(f"at most {L.ML_MAX_LAYERS - 1} distinct layers are supported")
    keep = []

    def scalar(value):
        v = float(_lowering._scalar(u.length(value), config_shape, cindex)) if value is not None else 0.0
        t = torch.full((1,), v, dtype=torch.float64, device=device)
        keep.append(t)
        return t

    inp = L.MlInput()
    inp.n_axes = 1
    inp.dims[0] = n_ray
    inp.wavelength = wavelength.data_ptr()
    inp.wavelength_stride[0] = 1
    inp.direction_re = cos_incidence.data_ptr()
    inp.direction_stride[0] = 1
    inp.n_re = index_refraction.data_ptr()
    inp.n_stride[0] = 1
    if bool((attenuation != 0).any().item()):
        k = attenuation * wavelength * (1.0 / (4 * np.pi))  # _multilayers.py:853
        keep.append(k)
        inp.n_im = k.data_ptr()

    constants = {}
    table = (L.MlLayer * len(stack))()
    for j, layer in enumerate(stack):
        chemical = layer._chemical if layer.chemical is not None else None
        if chemical is None:
            table[j].n_re = scalar(1.0).data_ptr()
        else:
            key = chemical.file_nk
            if key not in constants:
                xp, fr, fi, m = _nk_table(chemical, device, torch)
                re = torch.empty(n_ray, dtype=torch.float64, device=device)
                im = torch.empty(n_ray, dtype=torch.float64, device=device)
                L.check(
                    lib.optk_interp(
                        n_ray, wavelength.data_ptr(), m, xp.data_ptr(), fr.data_ptr(), fi.data_ptr(),
                        re.data_ptr(), im.data_ptr(), stream,
                    )
                )
                constants[key] = (re, im)
            re, im = constants[key]
            table[j].n_re, table[j].n_im = re.data_ptr(), im.data_ptr()
            table[j].n_stride[0] = 1
        table[j].thickness = scalar(layer.thickness).data_ptr()
        if layer.interface is not None:
            table[j].width = scalar(layer.interface.width).data_ptr()
            table[j].profile_kind = layer.interface.kind
    segs = (L.MlSegment * max(len(segments), 1))()
    for g, (first, count, repeat) in enumerate(segments):
        segs[g].first, segs[g].count, segs[g].repeat = first, count, repeat
    out = torch.empty((4, n_ray), dtype=torch.float64, device=device)
    ptrs = [out.data_ptr() + 8 * n_ray * q for q in range(4)]
    L.check(
        lib.optk_multilayer(C.byref(inp), len(stack), table, len(segments), segs, ptrs[0], ptrs[1], ptrs[2], ptrs[3], stream)
    )
    torch.cuda.current_stream(device).synchronize()  # `keep` and `constants` may be released now
    return out
