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


_hp_groups = {}


def _high_priority_group(group=None):
    """
    A NCCL process group over the same ranks whose kernels run on a HIGH-PRIORITY stream.  The trace
    kernel keeps every SM full, and a collective on torch's default (normal-priority) NCCL stream only
    gets onto the SMs when the trace grid drains: measured at N = 2 on cfg 5, 18 ms per 400 MB
    ``reduce_scatter`` under the trace against 0.6 ms at high priority.  Created once per parent group;
    ``None`` (use the parent) when the backend is not NCCL or the option is unavailable.
    """
    import torch.distributed as dist

    key = id(group) if group is not None else None
    if key in _hp_groups:
        return _hp_groups[key]
    made = None
    try:
        if dist.get_backend(group) == "nccl":
            options = dist.ProcessGroupNCCL.Options()
            options.is_high_priority_stream = True
            ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
            made = dist.new_group(ranks=ranks, backend="nccl", pg_options=options)
    except Exception:  # pragma: no cover
        made = None
    _hp_groups[key] = made
    return made


class ImagePipeline:
    """
    Reduction over the ranks and read-back of detector planes, one configuration at a time, on a side
    stream -- so configuration k is summed and copied while configuration k + 1 is traced.

    `image` is a FUSED :class:`~optika_b200._engine.DeviceImage` (``zeros(..., fused=True,
    pad_to=world)``): the fp64 planes of a configuration are one row of ``buffer_f64``, the counts one
    row of ``buffer_i64``.  :meth:`submit` is called right after the launches of configuration `c` were
    queued on the current stream.  Every rank ends up owning the SUM of one 1 / world slice of every row
    and copies it to the shared page-locked host buffer over its own PCIe link; :meth:`finish` waits for
    all ranks and returns the host planes (complete on every rank: the buffer is shared).

    How the slices are summed (`transport`):

    * ``"nccl"`` (default): ``reduce_scatter`` (one collective per dtype and configuration) on a
      HIGH-PRIORITY NCCL stream (:func:`_high_priority_group`), queued from the side stream;
    * ``"peer"``: every rank opens the other ranks' plane buffers through CUDA
      IPC and PULLS its slice of each with the copy engines over NVLink (``optk_memcpy_async`` into a
      staging buffer), then adds the ``world`` pieces with a small elementwise kernel.  Ordering between
      processes comes from interprocess CUDA events (recorded after the trace of a configuration,
      waited on by the pulling streams) plus one host-side barrier per exposure.  No SM is needed for
      the transfer, which matters here: the trace kernel keeps every SM full (3 CTAs x 256 threads x 80
      registers), and a NCCL kernel that wants 40 k registers per CTA only gets onto an SM when the
      trace grid drains -- measured at N = 2 on cfg 5: 18 ms per 400 MB configuration under the trace
      against 0.3 ms on an idle GPU;
    * gloo (the CPU tests) has neither: ``all_reduce`` and a slice.

    With one rank the pipeline is just an overlapped read-back.
    """

    def __init__(self, image, device=None, group=None, to_host: bool = True, local: bool = False,
                 transport: str | None = None, rezero: bool = False):
        import os
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
        # copies, adds and read-back must not queue behind the trace grid: highest priority
        self.side = torch.cuda.Stream(self.device, priority=-1) if self.cuda else None
        if transport is None:
            transport = os.environ.get("OPTK_REDUCE_TRANSPORT", "nccl")
        self.transport = transport if (self.world > 1 and self.cuda) else "none"
        self._nccl_group = None
        if self.transport == "nccl":
            self._nccl_group = _high_priority_group(group)
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
        # `rezero`: hand the planes of a configuration back ZEROED as soon as they have been summed and
        # copied (on the side stream, under the trace of the next configuration), so that a series of
        # exposures does not start each one with a multi-gigabyte memset on the critical path
        self.rezero = rezero
        self._pending = []
        self._host_group = None
        if self.transport == "peer":
            try:
                self._open_peers()
            except Exception as e:  # no peer access / IPC: NCCL carries the planes instead
                import warnings

                warnings.warn(f"peer-memory reduction unavailable ({e}); using NCCL reduce_scatter")
                self.transport = "nccl"
            # every rank must take the same route
            import torch.distributed as dist

            flag = [self.transport]
            gathered = [None] * self.world
            dist.all_gather_object(gathered, flag[0], group=group)
            if any(t != "peer" for t in gathered):
                self.transport = "nccl"

    # -- peer-memory transport ---------------------------------------------------------------
    def _open_peers(self):
        """Exchange CUDA IPC handles of the plane buffers and of one interprocess event per configuration."""
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor

        for peer in range(torch.cuda.device_count()):
            if peer != self.device.index and not torch.cuda.can_device_access_peer(self.device.index, peer):
                raise RuntimeError(f"device {self.device.index} has no peer access to device {peer}")
        self._host_group = dist.new_group(backend="gloo")  # host-side barriers that do not touch the GPU queues
        self._done = [torch.cuda.Event(enable_timing=False, interprocess=True) for _ in range(self.n_config)]
        for e in self._done:
            e.record(torch.cuda.current_stream(self.device))  # an IPC handle needs a recorded event
        mine = dict(
            f64=reduce_tensor(self.image.buffer_f64),
            i64=reduce_tensor(self.image.buffer_i64) if self.row_i else None,
            events=[e.ipc_handle() for e in self._done],
            device=self.device.index,
        )
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self._peer_f, self._peer_i, self._peer_done = {}, {}, {}
        from . import _lib as L

        for r, other in enumerate(everyone):
            if r == self.rank:
                continue
            with torch.cuda.device(self.device):  # direct NVLink access to the peer's memory, not a staged copy
                L.check(L.lib().optk_enable_peer_access(int(other["device"])))
            rebuild, args = other["f64"]
            self._peer_f[r] = rebuild(*args)
            if self.row_i:
                rebuild, args = other["i64"]
                self._peer_i[r] = rebuild(*args)
            self._peer_done[r] = [torch.cuda.Event.from_ipc_handle(self.device, h) for h in other["events"]]
        n_f, n_i = self.row_f // self.world, self.row_i // self.world
        self._stage_f = torch.empty((self.world - 1, n_f), dtype=torch.float64, device=self.device)
        self._stage_i = torch.empty((self.world - 1, n_i), dtype=torch.int64, device=self.device) if self.row_i else None
        dist.barrier(group=self._host_group)

    def _pull(self, c: int):
        """Side stream: sum this rank's slice of configuration `c` over all ranks into ``shard_*[c]``."""
        import torch
        from . import _lib as L

        lib, stream = L.lib(), self.side.cuda_stream
        rows = [(self.image.buffer_f64, self._peer_f, self._stage_f, self.shard_f)]
        if self.row_i:
            rows.append((self.image.buffer_i64, self._peer_i, self._stage_i, self.shard_i))
        peers = sorted(self._peer_f)
        self.side.wait_event(self._done[c])
        for r in peers:
            self.side.wait_event(self._peer_done[r][c])
        for local, remote, stage, shard in rows:
            n = shard.shape[1]
            lo = self.rank * n
            for j, r in enumerate(peers):
                src = remote[r][c, lo:lo + n]
                L.check(lib.optk_memcpy_async(stage[j].data_ptr(), src.data_ptr(), 8 * n, stream))
            torch.sum(stage, dim=0, out=shard[c])
            shard[c] += local[c, lo:lo + n]

    def _flush(self):
        """All configurations of this exposure are queued on every rank: start pulling."""
        import torch
        import torch.distributed as dist

        dist.barrier(group=self._host_group)  # every rank has RECORDED its events: waits now see this exposure
        with torch.cuda.stream(self.side):
            for c in self._pending:
                marks = None
                if self.timing:
                    marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    self.side.wait_event(self._done[c])
                    marks[0].record(self.side)
                self._pull(c)
                if marks:
                    marks[1].record(self.side)
                self._read_back(c, dict(f=self.shard_f[c], i=self.shard_i[c] if self.row_i else None))
                if marks:
                    marks[2].record(self.side)
                    self.events.append((c, *marks))
        self._pending = []

    def _read_back(self, c: int, sources: dict):
        if self.to_host:
            for kind, host in (("f", self.host_f), ("i", self.host_i if self.row_i else None)):
                if host is None or sources.get(kind) is None:
                    continue
                src = sources[kind]
                n = src.numel()
                host[c, self.rank * n:(self.rank + 1) * n].copy_(src, non_blocking=True)
        if self.rezero and self.transport != "peer":  # (peer: other ranks may still be reading these planes)
            self.image.buffer_f64[c].zero_()
            if self.row_i:
                self.image.buffer_i64[c].zero_()

    def _side(self):
        import contextlib
        import torch

        return torch.cuda.stream(self.side) if self.cuda else contextlib.nullcontext()

    def submit(self, c: int) -> None:
        import torch
        import torch.distributed as dist

        if self.transport == "peer":
            # the planes of configuration c are complete once everything queued so far has run
            self._done[c].record(torch.cuda.current_stream(self.device))
            self._pending.append(c)
            if len(self._pending) == self.n_config:
                self._flush()
            return
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
                    dist.reduce_scatter_tensor(
                        shard[c], buf[c], op=dist.ReduceOp.SUM,
                        group=self._nccl_group if self._nccl_group is not None else self.group,
                    )
                    sources[kind] = shard[c]
                else:  # gloo
                    dist.all_reduce(buf[c], op=dist.ReduceOp.SUM, group=self.group)
                    n = buf.shape[1] // self.world
                    sources[kind] = buf[c, self.rank * n:(self.rank + 1) * n]
            if marks:
                marks[1].record(self.side)
            self._read_back(c, sources)
            if marks:
                marks[2].record(self.side)
                self.events.append((c, *marks))

    def finish(self) -> dict:
        """Wait for the queued reductions and copies (all ranks); the host planes as NumPy views."""
        import torch.distributed as dist

        if self.transport == "peer" and self._pending:
            self._flush()
        if self.cuda:
            self.side.synchronize()
        if self.world > 1:
            # nobody may reuse (zero) its planes, or read the host buffer, before every rank is done
            dist.barrier(group=self._host_group if self._host_group is not None else self.group)
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
        if self.transport == "peer":
            import torch.distributed as dist

            if self.cuda:
                self.side.synchronize()
            dist.barrier(group=self._host_group)  # nobody unmaps while a peer may still read
            self._peer_f = self._peer_i = self._peer_done = None
            self.transport = "closed"
        if self.host is not None:
            self.host_f = self.host_i = None
            self.host.close()
            self.host = None
