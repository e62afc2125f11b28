#!/usr/bin/env python
"""
Stop solver: wall time of `SequentialSystem.rayfunction_stops`
(21 x 21 stop samples, what `_denormalize_grid` asks for) with the whole Newton iteration on
the device (`optk_solve_stops`) against the host iteration around device traces.
Run on the GPU box:  python tools/measure_stops.py > gpurun_out/stops.json
"""
import json
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

from optika_b200 import _stops
import configs


class HostNewton(_stops.DeviceBackend):
    solve = None


def wall_ms(fn, warmup=2, reps=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    results = []
    for name, system in (
        ("cfg1 newtonian", configs.newtonian(3, 3)),
        ("cfg3 toroidal VLS, 3 wavelengths", configs.toroidal_vls(3, 3, 3)),
        ("cfg5 misaligned telescope, 8 tilts", configs.misaligned_telescope(3, 3, 64, 8)),
    ):
        device = wall_ms(lambda: system.rayfunction_stops(21, 21))
        host = wall_ms(lambda: system.rayfunction_stops(21, 21, backend=HostNewton))
        results.append(dict(config=name, ms_device_newton=device, ms_host_newton_device_traces=host))
    print(json.dumps(dict(results=results), indent=1))


if __name__ == "__main__":
    main()
