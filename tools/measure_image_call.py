#!/usr/bin/env python
"""
Wall-clock of ONE user call ``SequentialSystem.image(scene)`` for the cfg 5 exposure (8 tilts x 4096 x 4096 field
cells x 10 x 8 pupil cells = 1.07e10 rays -> 8 detector images of 4096 x 4096), with where the time goes on the host.
Run on the GPU box:  python tools/measure_image_call.py > gpurun_out/image_call.json
"""
import cProfile
import json
import pathlib
import pstats
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch

from optika_b200 import named as na, units as u
import configs

axes = ("wavelength", "field_x", "field_y", "pupil_x", "pupil_y")
n_pixel = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
system = configs.telescope_4k(num_tilt=8, num_pixel=n_pixel)
half = float(np.arctan(0.5 * n_pixel * 15e-3 / 3200.0))
wavelength = na.ScalarArray(np.linspace(499 * u.nm, 501 * u.nm, 2), axes[0])
field = na.Cartesian2dVectorArray(
    na.ScalarArray(np.linspace(-half, half, n_pixel + 1), axes[1]), na.ScalarArray(np.linspace(-half, half, n_pixel + 1), axes[2])
)
pupil = na.Cartesian2dVectorArray(na.ScalarArray(np.linspace(-160, 160, 11), axes[3]), na.ScalarArray(np.linspace(-160, 160, 9), axes[4]))
from optika_b200.vectors import SpectralPositionalVectorArray

scene = na.FunctionArray(inputs=SpectralPositionalVectorArray(wavelength=wavelength, position=field), outputs=1e15)
out = {}
for label, noise in (("first call (compiles / loads kernels)", False), ("noise=False", False), ("noise=True", True)):
    torch.cuda.synchronize()
    profile = cProfile.Profile()
    t0 = time.perf_counter()
    profile.enable()
    image = system.image(scene, pupil=pupil, axis_wavelength=axes[0], axis_field=axes[1:3], axis_pupil=axes[3:5],
                         normalized_field=False, normalized_pupil=False, noise=noise)
    profile.disable()
    torch.cuda.synchronize()
    seconds = time.perf_counter() - t0
    stats = pstats.Stats(profile)
    top = sorted(stats.stats.items(), key=lambda kv: -kv[1][3])[:60]
    mine = [
        {"function": f"{pathlib.Path(k[0]).name}:{k[1]}:{k[2]}", "cumulative_s": round(v[3], 3), "calls": v[0]}
        for k, v in top if "optika_b200" in k[0] or "numpy" in k[0] and v[3] > 0.2
    ][:22]
    total = float(np.sum(na.as_named_array(image.outputs).ndarray))
    out[label] = dict(seconds=seconds, electrons_total=total, shape=dict(na.as_named_array(image.outputs).shape), host_profile=mine)
    del image
print(json.dumps(out, indent=1))
