#!/usr/bin/env python
"""
Fused trace + bin of the cfg-2 system over pupil strips (what the ranks of a multi-GPU run own):
    python tools/measure_strips.py [lo hi]     # with lo hi: launch that strip a few times (for ncu)
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

from optika_b200 import _engine, named as na
import configs
from measure_configs import time_ms


def strip_system(lo, hi, n=100):
    system = configs.spherical_grating(100, n, 1, 2048)
    system.grid_input.pupil = na.Cartesian2dVectorLinearSpace(
        start=na.Cartesian2dVectorArray(lo, -45.0), stop=na.Cartesian2dVectorArray(hi, 45.0),
        axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=n, centers=True,
    )
    return system


def main():
    device = torch.device("cuda", 0)
    edges = na.ScalarArray(np.array([1e-6, 1e-2]), "wavelength")
    strips = [(-45.0, 45.0), (-45.0, 0.0), (0.0, 45.0), (-45.0, -22.5), (-22.5, 0.0)]
    if len(sys.argv) > 2:
        strips = [(float(sys.argv[1]), float(sys.argv[2]))]
    results = []
    for lo, hi in strips:
        system = strip_system(lo, hi)
        ex, ey = system.sensor.pixel_edges()
        image = _engine.DeviceImage.zeros(edges.ndarray, ex, ey, device, moments=True, counts=True)
        ms = time_ms(lambda: system.image_rays(edges, image=image, device=device, **configs.PHYSICAL))
        counts = image.counts
        results.append(dict(strip=[lo, hi], ms_per_1e8_rays=ms, pixels_hit=int((counts > 0).sum().item()),
                            rays_binned_per_pass=int(counts.sum().item()) // 8, max_per_pixel=int(counts.max().item()) // 8))
    print(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
