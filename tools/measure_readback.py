#!/usr/bin/env python
"""
Device -> host read-back rate when all ranks copy at once (the tail of the strong-scaling exposures):
every rank copies `--mb` megabytes from its GPU into (a) its own page-locked buffer (torch pin_memory),
(b) its slice of the SHARED page-locked buffer (POSIX shared memory registered with CUDA in every process, what
ImagePipeline uses), (c) the same after every rank has first touched its own slice.  Run under torchrun.
"""
import argparse
import json
import os
import sys
import pathlib

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=400)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from optika_b200 import distributed

    n = args.mb * (1 << 20) // 8
    src = torch.rand(n, dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device)

    def timed(dst):
        rates = []
        for _ in range(args.reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            dst.copy_(src, non_blocking=True)
            e1.record(stream)
            torch.cuda.synchronize(device)
            rates.append(8 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        return float(sum(rates[1:]) / args.reps)

    results = {}
    private = torch.empty(n, dtype=torch.float64, pin_memory=True)
    results["private_pinned"] = timed(private)
    shared = distributed.SharedHostBuffer(8 * n * world)
    results["shared_registered"] = timed(shared.view(torch.float64, 8 * n * rank, n))
    shared.close()
    if world > 1:
        # first touch by the owner of each slice, then registration
        from multiprocessing import shared_memory, resource_tracker
        from optika_b200 import _lib as L

        name = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=8 * n * world)
            name = [shm.name]
        dist.broadcast_object_list(name, src=0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name[0])
            resource_tracker.unregister(shm._name, "shared_memory")
        whole = torch.frombuffer(shm.buf, dtype=torch.float64, count=n * world)
        whole[rank * n:(rank + 1) * n].zero_()  # this process faults its own pages in
        dist.barrier()
        L.check(L.lib().optk_host_register(whole.data_ptr(), 8 * n * world))
        dist.barrier()
        results["shared_first_touch_by_owner"] = timed(whole[rank * n:(rank + 1) * n])
        L.lib().optk_host_unregister(whole.data_ptr())
        dist.barrier()
        del whole
        shm.close()
        if rank == 0:
            shm.unlink()
    t = torch.tensor([results.get(k, 0.0) for k in ("private_pinned", "shared_registered", "shared_first_touch_by_owner")],
                     dtype=torch.float64, device=device)
    if world > 1:
        total = t.clone()
        dist.all_reduce(total)
        lowest = t.clone()
        dist.all_reduce(lowest, op=dist.ReduceOp.MIN)
    else:
        total = lowest = t
    if rank == 0:
        keys = ("private_pinned", "shared_registered", "shared_first_touch_by_owner")
        print(json.dumps(dict(
            n_gpus=world, megabytes_per_rank=args.mb,
            aggregate_gbytes_per_s=dict(zip(keys, [float(v) for v in total.tolist()])),
            slowest_rank_gbytes_per_s=dict(zip(keys, [float(v) for v in lowest.tolist()])),
        )))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
