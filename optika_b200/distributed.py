"""
Multi-GPU: one process per GPU, rays sharded by contiguous field / pupil slab.

The reference has no multi-process mode; rays are
independent, so the trace needs no exchange at all.  The only collective on the
path is the sum of the detector image planes after binning
(``optika/sensors/_sensors.py:155-161`` summed over ranks), done with
``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).
Integer hit counts reduce exactly; weighted fp64 sums to rounding.
"""

from __future__ import annotations
import dataclasses
import numpy as np
from . import named as na
from .vectors import ObjectVectorArray

__all__ = [
    "slab", "shard_axis", "shard_grid", "reduce_image", "image_rays_sharded", "rank_world", "best_shard_axis",
    "SharedHostBuffer", "ImagePipeline",
]


def rank_world() -> tuple[int, int]:
    """(rank, world size) of the default process group; (0, 1) outside ``torch.distributed``."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def slab(n: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slab of ``range(n)`` owned by `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} out of range for world size {world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def shard_axis(a, axis: str, rank: int, world: int):
    """The slab of named array (or vector) `a` along `axis`; objects without `axis` pass through."""
    if isinstance(a, (na.Cartesian2dVectorArray, na.Cartesian3dVectorArray)):
        return a._map(lambda c: shard_axis(c, axis, rank, world))
    if isinstance(a, na.ScalarArray) and axis in a.axes:
        return a[{axis: slab(a.shape[axis], rank, world)}]
    return a


def shard_grid(grid: ObjectVectorArray, axis: str, rank: int, world: int) -> ObjectVectorArray:
    """This rank's slab of an input grid along the named `axis` (e.g. ``"pupil_x"``)."""
    return ObjectVectorArray(
        wavelength=shard_axis(grid.wavelength, axis, rank, world),
        field=shard_axis(grid.field, axis, rank, world),
        pupil=shard_axis(grid.pupil, axis, rank, world),
    )


def reduce_image(planes, group=None):
    """
    Sum detector planes over all ranks, in place (``all_reduce``).  `planes` is a
    :class:`~optika_b200._engine.DeviceImage` or an iterable of tensors.
    """
    import torch.distributed as dist

    if hasattr(planes, "flux"):
        tensors = [t for t in (planes.flux, planes.moment_real, planes.moment_imag, planes.counts) if t is not None]
    else:
        tensors = list(planes)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return planes


def image_rays_sharded(system, wavelength_edges, axis: str = "pupil_x", grid=None, device=None, **kwargs):
    """
    Detector image of the whole input grid with the rays sharded along `axis`
    over the ranks of the default process group: every rank traces and bins its
    slab (fused kernel), then the planes are summed with one all-reduce.
    """
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    grid = system.grid_input if grid is None else grid
    mine = shard_grid(grid, axis, rank, world)
    image = system.image_rays(
        wavelength_edges, wavelength=mine.wavelength, field=mine.field, pupil=mine.pupil, device=device, **kwargs
    )
    return reduce_image(image)


def best_shard_axis(count, world: int, candidates=(3, 1, 4, 2)) -> int:
    """
    The grid axis to cut into `world` slabs: the one whose largest slab is closest to the mean
    (``ceil(n / world) / (n / world)`` minimal; an axis with fewer cells than ranks would leave ranks
    idle and scores infinitely badly), ties in the order of `candidates` (pupil x, field x, pupil y,
    field y: pupil slabs keep every rank on the whole detector, field slabs on a strip of it).
    """
    def imbalance(n):
        return float("inf") if n < world else -(-n // world) * world / n

    return min(candidates, key=lambda a: (imbalance(count[a]), candidates.index(a)))


class SharedHostBuffer:
    """
    One page-locked host buffer that every rank of a single-node job can write into: a POSIX
    shared-memory segment created by rank 0, mapped by the others and registered with CUDA in each
    process (``optk_host_register``).  Each rank copies ITS shard of the reduced detector planes
    device -> host over its own PCIe link; rank 0 then sees the whole image without any host-side
    gather.  Outside ``torch.distributed`` it is a plain pinned tensor.
    """

    def __init__(self, n_bytes: int, group=None, local: bool = False):
        import torch
        import torch.distributed as dist
        from . import _lib as L

        self.n_bytes = int(n_bytes)
        self.shm = None
        self._registered = None
        rank, world = (0, 1) if local else rank_world()
        self._rank = rank
        if world == 1:
            self.bytes = torch.empty(self.n_bytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            return
        from multiprocessing import shared_memory, resource_tracker
        import os

        name = [None]
        if rank == 0:
            free = os.statvfs("/dev/shm")
            if free.f_bavail * free.f_frsize < self.n_bytes + (64 << 20):
                name = ["!no space in /dev/shm"]
            else:
                self.shm = shared_memory.SharedMemory(create=True, size=self.n_bytes)
                name = [self.shm.name]
        dist.broadcast_object_list(name, src=0, group=group)
        if name[0].startswith("!"):
            raise MemoryError(f"SharedHostBuffer: {name[0][1:]} for {self.n_bytes} bytes")
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:  # the creator owns the segment: keep this process' resource tracker from unlinking it
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # pragma: no cover
                pass
        self.bytes = torch.frombuffer(self.shm.buf, dtype=torch.uint8, count=self.n_bytes)
        if torch.cuda.is_available():
            L.check(L.lib().optk_host_register(self.bytes.data_ptr(), self.n_bytes))
            self._registered = self.bytes.data_ptr()
        dist.barrier(group=group)

    def view(self, dtype, offset_bytes: int, count: int):
        import torch

        size = torch.empty((), dtype=dtype).element_size()
        return self.bytes[offset_bytes: offset_bytes + count * size].view(dtype)

    def close(self):
        from . import _lib as L

        if self._registered is not None:
            L.lib().optk_host_unregister(self._registered)
            self._registered = None
        self.bytes = None
        if self.shm is not None:
            try:
                self.shm.close()
                if self._rank == 0:
                    self.shm.unlink()
            except Exception:  # pragma: no cover
                pass
            self.shm = None


class ImagePipeline:
    """
    Reduction over the ranks and read-back of detector planes, one configuration at a time, on a side
    stream -- so configuration k is summed and copied while configuration k + 1 is traced.

    `image` is a FUSED :class:`~optika_b200._engine.DeviceImage` (``zeros(..., fused=True,
    pad_to=world)``): the fp64 planes of a configuration are one row of ``buffer_f64``, the counts one
    row of ``buffer_i64``.  :meth:`submit` (called right after the launches of configuration `c` were
    queued on the current stream) makes the side stream wait for them and then

    * ``reduce_scatter`` s the row over the ranks (one collective per dtype; NCCL over NVLink) --
      every rank ends up with the SUM of one 1 / world slice, which is all it needs because
    * each rank copies its slice to the shared page-locked host buffer over its own PCIe link.

    :meth:`finish` waits for the side stream and for the other ranks and returns the host planes
    (complete on every rank: the buffer is shared).  With one rank it is just an overlapped read-back.
    Gloo (the CPU tests) has no reduce-scatter: ``all_reduce`` and a slice take its place there.
    """

    def __init__(self, image, device=None, group=None, to_host: bool = True, local: bool = False):
        import torch

        if image.buffer_f64 is None:
            raise ValueError("ImagePipeline needs a fused DeviceImage (DeviceImage.zeros(..., fused=True))")
        self.image, self.group, self.to_host = image, group, to_host
        # `local`: this process alone (no collective, private host buffer) even inside a process group
        self.rank, self.world = (0, 1) if local else rank_world()
        self.cuda = image.buffer_f64.is_cuda
        self.device = image.buffer_f64.device
        n_config, row_f = image.buffer_f64.shape
        row_i = image.buffer_i64.shape[1] if image.buffer_i64 is not None else 0
        if row_f % self.world or row_i % self.world:
            raise ValueError(f"fused rows must be padded to a multiple of the world size (pad_to={self.world})")
        self.n_config, self.row_f, self.row_i = n_config, row_f, row_i
        self.side = torch.cuda.Stream(self.device) if self.cuda else None
        self.shard_f = torch.empty((n_config, row_f // self.world), dtype=torch.float64, device=self.device) \
            if self.world > 1 else None
        self.shard_i = torch.empty((n_config, row_i // self.world), dtype=torch.int64, device=self.device) \
            if self.world > 1 and row_i else None
        self.host = None
        if to_host:
            self.host = SharedHostBuffer(8 * n_config * (row_f + row_i), group=group, local=local)
            self.host_f = self.host.view(torch.float64, 0, n_config * row_f).view(n_config, row_f)
            self.host_i = self.host.view(torch.int64, 8 * n_config * row_f, n_config * row_i).view(n_config, row_i) \
                if row_i else None
        self.events = []  # (config, start, reduced, copied) CUDA events of the side stream
        self.timing = False

    def _side(self):
        import contextlib
        import torch

        return torch.cuda.stream(self.side) if self.cuda else contextlib.nullcontext()

    def submit(self, c: int) -> None:
        import torch
        import torch.distributed as dist

        if self.cuda:
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self.side.wait_event(done)
        with self._side():
            marks = None
            if self.cuda and self.timing:
                marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                marks[0].record(self.side)
            rows = [(self.image.buffer_f64, self.shard_f, "f")]
            if self.row_i:
                rows.append((self.image.buffer_i64, self.shard_i, "i"))
            sources = {}
            for buf, shard, kind in rows:
                if self.world == 1:
                    sources[kind] = buf[c]
                elif self.cuda:
                    dist.reduce_scatter_tensor(shard[c], buf[c], op=dist.ReduceOp.SUM, group=self.group)
                    sources[kind] = shard[c]
                else:  # gloo
                    dist.all_reduce(buf[c], op=dist.ReduceOp.SUM, group=self.group)
                    n = buf.shape[1] // self.world
                    sources[kind] = buf[c, self.rank * n:(self.rank + 1) * n]
            if marks:
                marks[1].record(self.side)
            if self.to_host:
                for kind, host in (("f", self.host_f), ("i", self.host_i if self.row_i else None)):
                    if host is None:
                        continue
                    src = sources[kind]
                    n = src.numel()
                    host[c, self.rank * n:(self.rank + 1) * n].copy_(src, non_blocking=True)
            if marks:
                marks[2].record(self.side)
                self.events.append((c, *marks))

    def finish(self) -> dict:
        """Wait for the queued reductions and copies (all ranks); the host planes as NumPy views."""
        import torch.distributed as dist

        if self.cuda:
            self.side.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        if not self.to_host:
            return {}
        im = self.image
        dims = tuple(im.flux.shape)
        n = int(np.prod(dims[-3:], dtype=np.int64))
        lead = dims[:-3]
        out = dict(flux=self.host_f[:, 0:n].numpy().reshape(lead + dims[-3:]))
        if im.moment_real is not None:
            out["moment_real"] = self.host_f[:, n:2 * n].numpy().reshape(lead + dims[-3:])
        if self.row_i:
            out["counts"] = self.host_i[:, 0:n].numpy().reshape(lead + dims[-3:])
        return out

    def stage_ms(self) -> dict:
        """Busy time of the side stream per stage, summed over the configurations (needs ``timing``)."""
        reduce_ms = sum(a.elapsed_time(b) for _, a, b, _ in self.events)
        copy_ms = sum(b.elapsed_time(c) for _, _, b, c in self.events)
        self.events = []
        return dict(ms_reduce=reduce_ms, ms_d2h=copy_ms)

    def close(self):
        if self.host is not None:
            self.host_f = self.host_i = None
            self.host.close()
            self.host = None
