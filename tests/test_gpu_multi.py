"""
Multi-GPU parity on hardware: the image of a simulation sharded over N ranks (one process per GPU,
torchrun, NCCL reduce_scatter + per-rank read-back into a shared host buffer) equals the image of one GPU
-- integer counts bit for bit, weighted sums to rounding.  Needs >= 2 GPUs.
"""

import json
import pathlib
import socket
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world,transport", [(2, "nccl"), (4, "nccl"), (8, "nccl"), (2, "peer")])
def test_sharded_reduced_image_equals_single_gpu_image(world, transport):
    import os
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    command = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
        str(ROOT / "tests" / "workers" / "multi_gpu_image.py"),
    ]
    # "nccl": reduce_scatter on a high-priority group (the default); "peer": CUDA IPC + copy engines + interprocess events
    env = dict(os.environ, OPTK_REDUCE_TRANSPORT=transport)
    done = subprocess.run(command, capture_output=True, text=True, timeout=900, cwd=str(ROOT), env=env)
    assert done.returncode == 0, done.stdout[-3000:] + done.stderr[-3000:]
    line = [ln for ln in done.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    report = json.loads(line[len("RESULT "):])
    assert set(report) == {"telescope", "toroidal_vls"}
    for name, r in report.items():
        assert r["binned"] > 0.3 * r["rays"], (name, r)
        assert r["counts_equal"] and r["counts_equal_one_call"], (name, r)
        assert r["flux_max_diff"] < 1e-12 and r["moment_max_diff"] < 1e-12, (name, r)
