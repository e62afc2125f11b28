#!/usr/bin/env python
"""
Throughput of the Monte-Carlo electron kernel (optk_electrons_measured) on a full detector: a 4096 x 4096 image with
Poisson(80) absorbed photons per pixel (the photon budget of the cfg 5 exposure), at 500 nm (one pair per photon),
30.4 nm (41 eV: ~11 pairs, collection-efficiency ramp) and 1 nm (1.24 keV: ~340 pairs).  CUDA events, 1 + 3 runs.
    python tools/measure_electrons.py > gpurun_out/electrons.json
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

from optika_b200 import named as na, sensors, units as u


def main():
    rng = np.random.default_rng(0)
    results = []
    for wavelength, n_pix, mean in ((500 * u.nm, 4096, 80), (30.4 * u.nm, 4096, 80), (1 * u.nm, 2048, 20)):
        photons = na.ScalarArray(rng.poisson(mean, size=(n_pix, n_pix)).astype(np.int64), ("x", "y"))
        kwargs = dict(thickness_implant=2000 * u.AA, thickness_depletion=5 * u.um, thickness_substrate=14 * u.um,
                      width_pixel=15 * u.um, cce_backsurface=0.6, axis_xy=("x", "y"))
        sensors.electrons_measured(photons, wavelength, seed=1, **kwargs)  # warm-up (tables, upload paths)
        torch.cuda.synchronize()
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            electrons = sensors.electrons_measured(photons, wavelength, seed=2 + rep, **kwargs)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        total_photons = int(photons.ndarray.sum())
        total_electrons = float(electrons.ndarray.sum())
        results.append(dict(
            wavelength_nm=wavelength / u.nm, pixels=n_pix * n_pix, photons=total_photons, electrons=total_electrons,
            electrons_per_photon=total_electrons / total_photons, seconds_end_to_end=best,
            photons_per_s=total_photons / best, electrons_per_s=total_electrons / best,
            note="wall clock of the public call: host -> device upload of the photon counts, kernel, device -> host counts",
        ))
    print(json.dumps(dict(results=results), indent=1))


if __name__ == "__main__":
    main()
