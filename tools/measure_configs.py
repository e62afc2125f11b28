#!/usr/bin/env python
"""
Device-side throughput of the BASELINE configurations that are NOT the bench line
(cfg 1, 3, 5 traces; cfg 4 multilayer sweep), for the tables in DESIGN.md.
Run on the GPU box:  python tools/measure_configs.py > gpurun_out/configs.json
Timing: CUDA events on the launch stream, 3 warm-up + 5 timed repetitions.
"""

import ctypes as C
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

import optika_b200 as optika
from optika_b200 import _engine, _lib, named as na, units as u
import configs


def time_ms(fn, warmup=3, reps=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def trace_config(name, system, n_surfaces, flop_per_ray):
    """Broadcast (separable) input -> dense output, and fused trace + bin."""
    device = torch.device("cuda", 0)
    _, rays = system._input(None, None, None, None, False, False)
    n = int(np.prod(list(rays.shape.values()))) * max(1, int(np.prod(list(system.shape.values()) or [1])))
    order = system._ray_axes_order
    ms_rays = time_ms(lambda: _engine.trace(system._compiled, rays, ray_axes_order=order, device=device))
    dense = _engine.trace(system._compiled_local, rays, surf_count=1, ray_axes_order=order, device=device)  # object surface only
    ms_dense = time_ms(lambda: _engine.trace(system._compiled, dense, device=device))
    edges = na.ScalarArray(np.array([1e-6, 1e-2]), "wavelength")
    ex, ey = system.sensor.pixel_edges()
    image = _engine.DeviceImage.zeros(
        edges.ndarray, ex, ey, device, leading=tuple(system._compiled.shape.values()), moments=True, counts=True
    )
    ms_image = time_ms(lambda: system.image_rays(edges, image=image, device=device, **configs.PHYSICAL))
    _, stats = _engine.trace(system._compiled, rays, ray_axes_order=order, device=device, stats=True)
    return dict(
        config=name,
        rays=n,
        surfaces=n_surfaces,
        ms_broadcast_in_dense_out=ms_rays,
        ms_dense_in_dense_out=ms_dense,
        ms_fused_trace_bin=ms_image,
        intercepts_per_s_dense=n * n_surfaces / (ms_dense * 1e-3),
        intercepts_per_s_fused_image=n * n_surfaces / (ms_image * 1e-3),
        hbm_gbs_dense=162.0 * n / (ms_dense * 1e-3) / 1e9,
        fp64_tflops_algorithmic_dense=flop_per_ray * n / (ms_dense * 1e-3) / 1e12,
        newton_iterations_per_ray=stats["n_newton_iterations"] / max(stats["n_rays"], 1),
        unvignetted_fraction=stats["n_unvignetted"] / max(stats["n_rays"], 1),
    )


def bin_config():
    """Kernel 2 alone: 1e8 traced cfg-2 rays resident in HBM -> 2048^2 image (41 B read per ray)."""
    device = torch.device("cuda", 0)
    system = configs.spherical_grating(100, 100, 1, 2048)
    _, rays = system._input(None, None, None, None, False, False)
    out = _engine.trace(system._compiled_local, rays, ray_axes_order=system._ray_axes_order, device=device)
    n = out.size
    ex, ey = system.sensor.pixel_edges()
    image = _engine.DeviceImage.zeros(np.array([1e-6, 1e-2]), ex, ey, device, moments=True, counts=True)
    im = image.struct(0)
    f = out.fields
    lib = _lib.lib()

    def run():
        _lib.check(
            lib.optk_bin(
                n, f["wavelength"].data_ptr(), f["px"].data_ptr(), f["py"].data_ptr(), f["dz"].data_ptr(),
                f["intensity"].data_ptr(), out.unvignetted.data_ptr(), C.byref(im), None,
            )
        )

    ms = time_ms(run)
    return dict(
        config="kernel 2 (optk_bin), cfg2 rays at the sensor", rays=n, ms=ms, rays_per_s=n / (ms * 1e-3),
        hbm_gbs=41.0 * n / (ms * 1e-3) / 1e9, binned_fraction=float(image.counts.sum().item()) / (8.0 * n),
    )


def reduce_config():
    """optk_reduce_groups: 1e8 cfg-1 rays at the sensor reduced over the pupil (25 B read per ray)."""
    device = torch.device("cuda", 0)
    system = configs.newtonian(100, 100, 128)
    _, rays = system._input(None, None, None, None, False, False)
    out = _engine.trace(system._compiled_local, rays, ray_axes_order=system._ray_axes_order, device=device)
    ms = time_ms(lambda: _engine.reduce_groups(out, ("pupil_x", "pupil_y"), device=device))
    # fused: trace + reduce in one launch, no ray written (optk_image_t.group_size)
    groups = _engine.DeviceGroups.zeros(1, 100 * 100, 100 * 100, device)
    ms_fused = time_ms(lambda: _engine.trace(system._compiled_local, rays, image=groups, write_rays=False,
                                             ray_axes_order=system._ray_axes_order, device=device))
    return dict(
        ms_fused_trace_reduce=ms_fused, intercepts_per_s_fused_trace_reduce=out.size * 6 / (ms_fused * 1e-3),
        config="optk_reduce_groups, cfg1 rays at the sensor, 1e4 field points x 1e4 pupil samples", rays=out.size, ms=ms,
        rays_per_s=out.size / (ms * 1e-3), hbm_gbs=25.0 * out.size / (ms * 1e-3) / 1e9,
    )


def multilayer_config(n_w=4096, n_t=1024, n_c=256, bilayers=30):
    """cfg 4: 60-layer Mo/Si stack on SiO2, erf interfaces, thickness scaled per configuration."""
    device = torch.device("cuda", 0)
    M = optika.materials
    w = na.linspace(10 * u.nm, 15 * u.nm, axis="wavelength", num=n_w)
    cos = na.linspace(np.cos(np.deg2rad(30)), 1.0, axis="angle", num=n_t)
    scale = na.linspace(0.95, 1.05, axis="config", num=n_c)
    d, gamma = 6.65 * u.nm, 0.6
    si = M.Layer("Si", thickness=scale * d * gamma, interface=M.profiles.ErfInterfaceProfile(0.7 * u.nm))
    mo = M.Layer("Mo", thickness=scale * d * (1 - gamma), interface=M.profiles.ErfInterfaceProfile(0.7 * u.nm))
    explicit = M.LayerSequence([si, mo] * bilayers)
    periodic = M.PeriodicLayerSequence([si, mo], num_periods=bilayers)
    substrate = M.Layer("SiO2", interface=M.profiles.ErfInterfaceProfile(0.7 * u.nm))
    out = {}
    for name, layers in (("explicit_60_layers", explicit), ("periodic_30x2", periodic)):
        fn = lambda: M._multilayers.multilayer_efficiency_device(w, cos, 1, layers, substrate, device=device)  # noqa: E731
        ms = time_ms(fn, warmup=1, reps=2)
        n = n_w * n_t * n_c
        out[name] = dict(
            evaluations=n,
            ms=ms,
            evaluations_per_s=n / (ms * 1e-3),
            fp64_tflops_algorithmic=2.4e4 * n / (ms * 1e-3) / 1e12,
        )
    return dict(config="cfg4 multilayer", grid=[n_w, n_t, n_c], **out)


def main():
    results = []
    results.append(trace_config("cfg1 newtonian 100x100 field x 100x100 pupil", configs.newtonian(100, 100, 128), 6, 791))
    results.append(trace_config("cfg3 toroidal VLS + octagon 100x100x100x100 x 8 wl/8", configs.toroidal_vls(100, 100, 1), 4, 771))
    results.append(
        trace_config("cfg5 misaligned telescope 8 tilts x 64x64 field x 100x100 pupil, 4096^2 sensor",
                     configs.misaligned_telescope(64, 100, 4096, 8), 6, 791)
    )
    results.append(bin_config())
    results.append(reduce_config())
    results.append(multilayer_config())
    fp64 = C.c_double()
    _lib.check(_lib.lib().optk_measure_fp64_peak(C.byref(fp64), None))
    soa = C.c_double()
    _lib.check(_lib.lib().optk_measure_soa_copy(100_000_000, C.byref(soa), None))
    print(json.dumps(dict(fp64_peak_tflops=fp64.value / 1e12, soa_copy_gbs=soa.value, results=results), indent=1))


if __name__ == "__main__":
    main()
