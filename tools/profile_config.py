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
    }[name]()
    device = torch.device("cuda", 0)
    _, rays = system._input(None, None, None, None, False, False)
    order = system._ray_axes_order
    if mode == "dense":
        dense = _engine.trace(system._compiled_local, rays, surf_count=1, ray_axes_order=order, device=device)
        for _ in range(4):
            _engine.trace(system._compiled, dense, device=device)
    else:
        edges = na.ScalarArray(np.array([1e-6, 1e-2]), "wavelength")
        ex, ey = system.sensor.pixel_edges()
        image = _engine.DeviceImage.zeros(edges.ndarray, ex, ey, device, leading=tuple(system._compiled.shape.values()),
                                          moments=True, counts=True)
        for _ in range(4):
            system.image_rays(edges, image=image, device=device)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
