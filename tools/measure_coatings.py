#!/usr/bin/env python
"""
Multilayer-coated cfg 2 grating (periodic Mo/Si stack on the concave grating): fused trace + detector image of
1e8 rays with the coating evaluated exactly per ray (chained launches) and from an efficiency table inside the
fused launch (optika_b200/_coatings.py).  Run on the GPU box:  python tools/measure_coatings.py > gpurun_out/coatings.json
Timing: CUDA events, 1 warm-up (table construction, kernel compilation) + 3 timed repetitions.
"""

import json
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

import optika_b200 as optika
from optika_b200 import _engine, named as na, units as u
import configs
from measure_configs import time_ms

M = optika.materials


def mo_si(num_periods):
    d, gamma = 6.65 * u.nm, 0.6
    rough = M.profiles.ErfInterfaceProfile(0.7 * u.nm)
    return M.MultilayerMirror(
        layers=[
            M.Layer("SiO2", thickness=1 * u.nm),
            M.PeriodicLayerSequence(
                [M.Layer("Si", thickness=d * gamma, interface=rough), M.Layer("Mo", thickness=d * (1 - gamma), interface=rough)],
                num_periods=num_periods,
            ),
        ],
        substrate=M.Layer("SiO2", interface=rough),
    )


def system_for(num_periods, num_field, num_pupil, num_wavelength):
    system = configs.spherical_grating(num_field=num_field, num_pupil=num_pupil, num_wavelength=num_wavelength, num_pixel=2048)
    system.grid_input.wavelength = na.linspace(12.5 * u.nm, 14.5 * u.nm, axis="wavelength", num=num_wavelength)
    grating = system.surfaces[0]
    grating.material = mo_si(num_periods)
    grating.rulings = optika.rulings.Rulings(spacing=(13.5 / 40 / 1200) * u.mm, diffraction_order=1)
    return system


def main():
    device = torch.device("cuda", 0)
    results = []
    edges = na.ScalarArray(np.array([12 * u.nm, 15 * u.nm]), "wavelength")
    for num_periods in (10, 30):
        system = system_for(num_periods, 100, 100, 1)  # 1e8 rays, one wavelength
        ex, ey = system.sensor.pixel_edges()
        row = dict(config=f"cfg2 grating with a periodic Mo/Si coating, {num_periods} periods, 1e8 rays", rays=10**8)
        images = {}
        for mode in ("exact", "table"):
            system.coating = mode
            image = _engine.DeviceImage.zeros(edges.ndarray, ex, ey, device, moments=True, counts=True)
            t0 = time.perf_counter()
            system.image_rays(edges, image=image, device=device, **configs.PHYSICAL)
            torch.cuda.synchronize()
            row[f"first_call_s_{mode}"] = time.perf_counter() - t0
            image.zero_()
            row[f"ms_{mode}"] = time_ms(lambda: system.image_rays(edges, image=image, device=device, **configs.PHYSICAL), warmup=1, reps=3)
            images[mode] = image.flux.clone() / 4.0  # 1 + 3 repetitions accumulated
        a, b = images["table"], images["exact"]
        row["flux_max_abs_diff_over_max"] = float((a - b).abs().max().item() / b.max().item())
        tabled = [t for t in system._compiled_local.__dict__.get("_tabled", {}).values() if t != "exact"]
        if tabled:
            t = next(iter(tabled[0].tables.values()))
            row["table"] = dict(n_wavelength=t.n_w, n_cos=t.n_c, bytes=t.nbytes, cos_range=t.cos_range, error_cos=t.error_cos)
        row["speedup"] = row["ms_exact"] / row["ms_table"]
        results.append(row)
    print(json.dumps(dict(results=results), indent=1))


if __name__ == "__main__":
    main()
