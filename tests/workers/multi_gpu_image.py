"""
Worker of tests/test_gpu_multi.py, launched under torchrun with one rank per GPU: the detector image of a
sharded, reduced (NCCL reduce_scatter) and read-back simulation must equal the image one GPU computes
alone -- counts bit for bit, weighted sums to rounding.  Prints one JSON line on rank 0.
"""

import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[2]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    import configs
    from optika_b200 import _engine, distributed, named as na, units as u

    report = {}
    axes = ("wavelength", "field_x", "field_y", "pupil_x", "pupil_y")
    cases = {
        # configuration axis (3 tilts), field-sharded (the pupil axes are too short for the world size)
        "telescope": (
            configs.telescope_4k(num_tilt=3, num_pixel=256),
            (na.ScalarArray(np.linspace(499e-6, 501e-6, 2), axes[0]),
             na.Cartesian2dVectorArray(na.ScalarArray(np.linspace(-3e-4, 3e-4, 97), axes[1]),
                                       na.ScalarArray(np.linspace(-3e-4, 3e-4, 65), axes[2])),
             na.Cartesian2dVectorArray(na.ScalarArray(np.linspace(-160, 160, 2), axes[3]),
                                       na.ScalarArray(np.linspace(-160, 160, 12), axes[4]))),
        ),
        # pupil-sharded, several wavelength cells, vignetting by an octagon
        "toroidal_vls": (
            configs.toroidal_vls(6, 12, 3),
            (na.ScalarArray(np.linspace(25e-6, 35e-6, 4), axes[0]),
             na.Cartesian2dVectorArray(na.ScalarArray(np.linspace(-0.2, 0.2, 11) * u.deg, axes[1]),
                                       na.ScalarArray(np.linspace(-0.2, 0.2, 9) * u.deg, axes[2])),
             na.Cartesian2dVectorArray(na.ScalarArray(np.linspace(-22, 22, 65), axes[3]),
                                       na.ScalarArray(np.linspace(-22, 22, 33), axes[4]))),
        ),
    }
    for name, (system, (wavelength, field, pupil)) in cases.items():
        grids = system.ray_grids(1.0, wavelength, field, pupil, axes[0], axes[1:3], axes[3:5],
                                 normalized_field=False, normalized_pupil=False, seed=3)
        w_edges = np.array([wavelength.ndarray.min(), wavelength.ndarray.max()])
        ex, ey = system.sensor.pixel_edges()
        leading = tuple(system._compiled_local.shape.values())
        image = _engine.DeviceImage.zeros(w_edges, ex, ey, device, leading=leading, counts=True, fused=True, pad_to=world)
        pipeline = distributed.ImagePipeline(image, device)
        for exposure in range(2):  # the second exposure reuses every buffer
            image.zero_()
            planes = system.collect_grids(grids, w_edges, device=device, pipeline=pipeline)
        sharded = {k: np.array(v) for k, v in planes.items()}
        dist.barrier()
        pipeline.close()
        # the public one-call form allocates its own buffers and must agree with the explicit pipeline
        again = system.collect_grids(grids, w_edges, device=device, counts=True)
        if rank == 0:
            whole = system.collect_grids(grids, w_edges, device=device, counts=True, shard=False, reduce=False)
            scale = float(np.abs(whole["flux"]).max())
            report[name] = dict(
                rays=int(sum(g.size for g in grids)),
                binned=int(whole["counts"].sum()),
                counts_equal=bool(np.array_equal(sharded["counts"], whole["counts"])),
                counts_equal_one_call=bool(np.array_equal(again["counts"], whole["counts"])),
                flux_max_diff=float(np.abs(sharded["flux"] - whole["flux"]).max() / scale),
                moment_max_diff=float(np.abs(sharded["moment_real"] - whole["moment_real"]).max() / scale),
                shard_axis=distributed.best_shard_axis(grids[0].count, world),
            )
        dist.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
