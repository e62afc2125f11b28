#!/usr/bin/env python
"""Where do the table-driven and the run-time compiled kernels stop being bit-identical?  (diagnostic, GPU)"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch

from optika_b200 import _engine, _lib
import configs

lib = _lib.lib()
system = configs.newtonian(num_field=5, num_pupil=12, num_pixel=64)
_, rays = system._input(None, None, None, None, False, False)
order = system._ray_axes_order
surfaces = system.surfaces_all
for k in range(1, len(surfaces) + 1):
    compiled = _engine.CompiledSystem(surfaces[:k])
    lib.optk_jit_mode(0)
    a = _engine.trace(compiled, rays, ray_axes_order=order)
    lib.optk_jit_mode(1)
    b = _engine.trace(compiled, rays, ray_axes_order=order)
    lib.optk_jit_mode(-1)
    line = []
    for name in a.fields:
        x, y = a.fields[name].reshape(-1), b.fields[name].reshape(-1)
        ne = (x != y) & ~(torch.isnan(x) & torch.isnan(y))
        if ne.any():
            rel = ((x - y).abs() / y.abs().clamp(min=1e-300))[ne].max().item()
            line.append(f"{name}: {int(ne.sum())} differ, max rel {rel:.2e}")
    print(k, type(surfaces[k - 1].sag).__name__ if surfaces[k - 1].sag is not None else None, hex(0), "; ".join(line) or "identical")
    if k == 4:
        x, y = a.fields["dx"].reshape(-1), b.fields["dx"].reshape(-1)
        idx = torch.nonzero(x != y).reshape(-1).tolist()
        print("   differing flat indices", idx, "of", x.numel(), "shape", tuple(a.fields["dx"].shape))
        for name in ("px", "py", "pz", "dx", "dy", "dz", "intensity", "wavelength"):
            if name in a.fields:
                print("   ", name, [float(a.fields[name].reshape(-1)[i]) for i in idx[:3]], [float(b.fields[name].reshape(-1)[i]) for i in idx[:3]])
