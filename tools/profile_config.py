#!/usr/bin/env python
"""
Launch one configuration's device-resident trace a few times, for an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:optk_jit_kernel -s 2 -c 1 \
      -o gpurun_out/prof python tools/profile_config.py cfg3 dense
mode: dense (dense SoA in -> dense SoA out), image (fused trace + bin from broadcast grids).
"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

from optika_b200 import _engine, named as na
import configs


def main():
    name, mode = sys.argv[1], sys.argv[2]
    system = {
        "cfg1": lambda: configs.newtonian(100, 100, 128),
        "cfg2": lambda: configs.spherical_grating(100, 100, 1, 2048),
        "cfg3": lambda: configs.toroidal_vls(100, 100, 1),
        "cfg5": lambda: configs.telescope_4k(num_tilt=1, num_pixel=4096),  # the round-2 full-detector telescope, one tilt
    }[name]()
    device = torch.device("cuda", 0)
    _, rays = system._input(None, None, None, None, False, False)
    order = system._ray_axes_order
    if mode == "dense":
        dense = _engine.trace(system._compiled_local, rays, surf_count=1, ray_axes_order=order, device=device)
        for _ in range(4):
            _engine.trace(system._compiled, dense, device=device)
    elif mode == "grid":
        # generate (Philox) + trace + bin: optk_trace_grid on a 1 x 100 x 100 x 100 x 100 vertex grid
        from optika_b200 import _grid, units as u

        deg = u.deg
        lin = lambda lo, hi, n: np.linspace(lo, hi, n + 1)  # noqa: E731
        vertices = {
            "cfg1": [lin(499 * u.nm, 501 * u.nm, 1), lin(-0.1 * deg, 0.1 * deg, 100), lin(-0.1 * deg, 0.1 * deg, 100), lin(-40, 40, 100), lin(-40, 40, 100)],
            "cfg2": [lin(17 * u.nm, 63 * u.nm, 1), lin(-0.05 * deg, 0.05 * deg, 100), lin(-0.05 * deg, 0.05 * deg, 100), lin(-45, 45, 100), lin(-45, 45, 100)],
            "cfg3": [lin(25 * u.nm, 35 * u.nm, 1), lin(-0.2 * deg, 0.2 * deg, 100), lin(-0.2 * deg, 0.2 * deg, 100), lin(-22, 22, 100), lin(-22, 22, 100)],
            # a quarter of the detector: 1118 x 1118 field cells (one per pixel) x 10 x 8 pupil cells = 1e8 rays
            "cfg5": [lin(499 * u.nm, 501 * u.nm, 1), lin(-0.0048, 0.0, 1118), lin(-0.0048, 0.0, 1118), lin(-160, 160, 10), lin(-160, 160, 8)],
        }[name]
        compiled = system._compiled_local
        ex, ey = system.sensor.pixel_edges()
        ew = np.array([vertices[0][0], vertices[0][-1]])
        image = _engine.DeviceImage.zeros(ew, ex, ey, device, leading=tuple(compiled.shape.values()), moments=True, counts=True)
        grid = _grid.RayGrid(vertices, jitter=True, seed=0)
        for _ in range(4):
            _grid.trace_grid(compiled, grid, config=0, image=image, write_rays=False, device=device)
    else:
        edges = na.ScalarArray(np.array([1e-6, 1e-2]), "wavelength")
        ex, ey = system.sensor.pixel_edges()
        image = _engine.DeviceImage.zeros(edges.ndarray, ex, ey, device, leading=tuple(system._compiled.shape.values()),
                                          moments=True, counts=True)
        for _ in range(4):
            system.image_rays(edges, image=image, device=device, **configs.PHYSICAL)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
