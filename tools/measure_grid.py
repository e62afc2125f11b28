#!/usr/bin/env python
"""
Device-side throughput of the on-device ray generator (optk_trace_grid): rays drawn,
traced and binned without being read from (or written to) HBM, for the BASELINE
configurations whose volumes only fit that way (cfg 2 "generated on chip", cfg 5).
Run on the GPU box:  python tools/measure_grid.py > gpurun_out/grid.json
Timing: CUDA events on the launch stream, 3 warm-up + 5 timed repetitions.
"""

import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

from optika_b200 import _engine, _grid, units as u
import configs
from measure_configs import time_ms


def measure(name, system, vertices, n_surfaces, flop_per_ray, jitter=True, rays_out=True):
    device = torch.device("cuda", 0)
    compiled = system._compiled_local
    ex, ey = system.sensor.pixel_edges()
    ew = np.array([vertices[0][0], vertices[0][-1]])
    n_config = compiled.n_config
    image = _engine.DeviceImage.zeros(
        ew, ex, ey, device, leading=tuple(compiled.shape.values()), moments=True, counts=True
    )
    grid = _grid.RayGrid(vertices, jitter=jitter, seed=0)
    n = grid.size * n_config

    def fused():
        for c in range(n_config):
            _grid.trace_grid(compiled, grid, config=c, image=image, write_rays=False, device=device)

    ms_image = time_ms(fused)
    _, stats = _grid.trace_grid(compiled, grid, config=0, write_rays=False, stats=True, device=device)
    out = dict(
        config=name,
        rays=n,
        surfaces=n_surfaces,
        jitter=jitter,
        ms_grid_fused_trace_bin=ms_image,
        intercepts_per_s_grid_fused=n * n_surfaces / (ms_image * 1e-3),
        fp64_tflops_algorithmic=flop_per_ray * n / (ms_image * 1e-3) / 1e12,
        unvignetted_fraction=stats["n_unvignetted"] / max(stats["n_rays"], 1),
        binned_fraction=float(image.counts.sum().item()) / (n * 9),  # 1 + 3 warm-up + 5 timed passes
    )
    if rays_out and grid.size * 81 * 1.2 < 60e9:
        ms_rays = time_ms(lambda: _grid.trace_grid(compiled, grid, config=0, device=device), warmup=2, reps=3)
        out["ms_grid_rays_out_one_config"] = ms_rays
        out["intercepts_per_s_grid_rays_out"] = grid.size * n_surfaces / (ms_rays * 1e-3)
        out["hbm_gbs_rays_out"] = 81.0 * grid.size / (ms_rays * 1e-3) / 1e9
    return out


def edges(lo, hi, n):
    return np.linspace(lo, hi, n + 1)


def main():
    results = []
    deg = u.deg
    for jitter in (True, False):
        results.append(measure(
            "cfg2 spherical grating, 100x100 field x 100x100 pupil x 4 wavelength cells",
            configs.spherical_grating(8, 16, 4, 2048),
            [edges(17 * u.nm, 63 * u.nm, 4), edges(-0.05 * deg, 0.05 * deg, 100), edges(-0.05 * deg, 0.05 * deg, 100),
             edges(-45, 45, 100), edges(-45, 45, 100)],
            3, 410, jitter=jitter,
        ))
    results.append(measure(
        "cfg1 newtonian, 100x100 field x 100x100 pupil",
        configs.newtonian(10, 32, 128),
        [edges(499 * u.nm, 501 * u.nm, 1), edges(-0.1 * deg, 0.1 * deg, 100), edges(-0.1 * deg, 0.1 * deg, 100),
         edges(-40, 40, 100), edges(-40, 40, 100)],
        6, 791,
    ))
    results.append(measure(
        "cfg5 misaligned telescope, 8 tilts x 354x354 field x 100x100 pupil = 1.0e10 rays, 4096^2 sensor",
        configs.misaligned_telescope(6, 12, 4096, 8),
        [edges(499 * u.nm, 501 * u.nm, 1), edges(-0.1 * deg, 0.1 * deg, 354), edges(-0.1 * deg, 0.1 * deg, 354),
         edges(-40, 40, 100), edges(-40, 40, 100)],
        6, 791, rays_out=False,
    ))
    results.append(measure(
        "cfg3 toroidal VLS, 100x100 field x 100x100 pupil x 1 wavelength cell",
        configs.toroidal_vls(6, 12, 3),
        [edges(25 * u.nm, 35 * u.nm, 1), edges(-0.2 * deg, 0.2 * deg, 100), edges(-0.2 * deg, 0.2 * deg, 100),
         edges(-22, 22, 100), edges(-22, 22, 100)],
        4, 771,
    ))
    print(json.dumps(dict(results=results), indent=1))


if __name__ == "__main__":
    main()
