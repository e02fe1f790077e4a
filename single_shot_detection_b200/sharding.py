"""Multi-GPU: shard the batch by image, one exchange at the end.

Every stage of the path is per-image independent and the anchors are a replicated constant, so
rank r simply runs the pipeline on images [r*B/W, (r+1)*B/W).  The only collective is a final
all-gather of the padded detections, their counts and the matched-target statistics, packed into
ONE buffer per rank (<= 4.8 KB per image) so that a single ``all_gather_into_tensor`` over
NCCL / NVLink moves everything.  The reference has no equivalent (its eval is replicated on every
rank, bf/builders/data_builder.py:58); the invariant is gathered result == single-GPU result.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def image_shard(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the images rank ``rank`` owns; shards differ by at most one image."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_capacity(batch: int, world: int) -> int:
    return (batch + world - 1) // world


def row_words(max_rows: int) -> int:
    """Words per packed row: T*6 detections, count, 4 statistics, zero padding to a 16-byte multiple
    (``ssd_shard_row_words`` of the C ABI)."""
    return (max_rows * 6 + 5 + 3) & ~3


def pack_shard(dets: torch.Tensor, counts: torch.Tensor, stats: torch.Tensor, capacity: int) -> torch.Tensor:
    """[capacity, row_words(T)] fp32 words: detections, then count and stats bit-cast from int32, then padding.
    Rows beyond the local image count have count = -1."""
    n, t = dets.shape[0], dets.shape[1]
    buf = torch.zeros((capacity, row_words(t)), dtype=torch.float32, device=dets.device)
    ints = buf.view(torch.int32)
    ints[:, t * 6] = -1
    if n:
        buf[:n, : t * 6] = dets.reshape(n, t * 6)
        ints[:n, t * 6] = counts.to(torch.int32)
        ints[:n, t * 6 + 1: t * 6 + 5] = stats.to(torch.int32)
    return buf


def pack_shard_device(dets: torch.Tensor, counts: torch.Tensor, assign_stats: Optional[torch.Tensor],
                      mining_stats: Optional[torch.Tensor], capacity: int):
    """:func:`pack_shard` + ``pipeline.matched_stats`` for CUDA tensors in ONE launch (csrc/exchange.cu):
    -> (shard [capacity, row_words(T)], stats [B, 4] int32)."""
    from . import _native as N
    N.require_device()
    n, t = int(dets.shape[0]), int(dets.shape[1])
    shard = torch.empty((capacity, row_words(t)), dtype=torch.float32, device=dets.device)
    stats = torch.empty((n, 4), dtype=torch.int32, device=dets.device)
    with torch.cuda.device(dets.device):
        N.check(N.lib().ssd_pack_shard(dets.data_ptr(), counts.data_ptr(),
                                       None if assign_stats is None else assign_stats.data_ptr(),
                                       None if mining_stats is None else mining_stats.data_ptr(), n, t, capacity,
                                       shard.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return shard, stats


class _DeviceBytes:
    """``__cuda_array_interface__`` over raw device memory, so that torch can alias it (torch.as_tensor)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2,
                                         "strides": None}


class PeerExchange:
    """The exchange step WITHOUT a collective library: every rank owns one arena of gathered buffers, maps the
    arenas of its peers through CUDA IPC (all ranks sit on one NVSwitch box) and ``pack_exchange`` -- one kernel,
    capturable into the step graph -- packs the local shard and writes it straight into every rank's buffer over
    NVLink (csrc/exchange.cu).  ``torch.distributed`` is used once, to hand the 64-byte IPC handles around; with a
    world of one (or without a process group) the same kernels run on the own arena alone.

    One ``slot`` per concurrently usable step graph; launches on the same slot must be serialised (same stream)
    and every launch is ``open(slot)`` ... ``pack_exchange(slot)``: ``open`` tells every rank that this one has
    finished with the slot's previous contents, and a writer waits until every rank has opened the same launch, so
    a rank that runs ahead never overwrites rows a slower peer is still reading.  ``gathered(slot)`` is complete
    once ``wait(slot)`` has run on the stream, and stays valid until the slot is opened again."""

    def __init__(self, batch: int, max_rows: int, slots: int, group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None):
        import ctypes
        from . import _native as N
        N.require_device()
        self.group = group
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if distributed else 1
        self.rank = dist.get_rank(group) if distributed else 0
        self.batch, self.max_rows, self.slots = batch, max_rows, slots
        self.capacity = shard_capacity(batch, self.world)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._N = N
        self._closed = False
        lib = N.lib()
        nbytes = lib.ssd_exchange_arena_bytes(self.world, slots, self.capacity, max_rows)
        if nbytes == 0:
            raise ValueError(f"PeerExchange: world {self.world} / slots {slots} / {self.capacity} images per rank "
                             "outside the supported range")
        self._nbytes = nbytes
        handle = ctypes.create_string_buffer(64)
        own = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            N.check(lib.ssd_exchange_arena_alloc(nbytes, ctypes.byref(own), handle))     # zero-filled, synchronised
        self._own = own.value
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, (bytes(handle.raw), int(self.device.index)), group=group)
        self._mapped = []                                   # peers' arenas as mapped here: closed in close()
        ptrs = (ctypes.c_void_p * self.world)()
        with torch.cuda.device(self.device):
            for r in range(self.world):
                if r == self.rank:
                    ptrs[r] = self._own
                    continue
                raw, peer_device = handles[r]
                # the kernels that use the mapping run on THIS device: the handle is opened with this device
                # current (cudaIpcOpenMemHandle enables peer access to the exporter's device lazily)
                N.check(lib.ssd_exchange_enable_peer(self.device.index, peer_device))
                mapped = ctypes.c_void_p()
                N.check(lib.ssd_exchange_peer_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(mapped)))
                self._mapped.append(mapped.value)
                ptrs[r] = mapped.value
        self._ptrs = ptrs
        self.arena = torch.as_tensor(_DeviceBytes(self._own, nbytes), device=self.device)   # aliases the arena
        if self.world > 1:
            dist.barrier(group=group)                      # every rank has mapped every arena

    def close(self) -> None:
        """Unmap the peers' arenas and free the own one (collective: every rank calls it before it exits, so that
        no exporter frees memory that is still mapped elsewhere)."""
        if self._closed:
            return
        self._closed = True
        lib = self._N.lib()
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for mapped in self._mapped:
                self._N.check(lib.ssd_exchange_peer_close(mapped))
            self._mapped = []
            if self.world > 1:
                dist.barrier(group=self.group)             # nobody maps this rank's arena any more
            self.arena = None
            self._N.check(lib.ssd_exchange_arena_free(self._own))
            self._own = None

    def _stream(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def open(self, slot: int) -> None:
        """First operation of a launch of ``slot`` (see the class docstring); capturable."""
        assert 0 <= slot < self.slots
        with torch.cuda.device(self.device):
            self._N.check(self._N.lib().ssd_exchange_open(self._ptrs, self.world, self.rank, slot, self._stream()))

    def pack_exchange(self, dets: torch.Tensor, counts: torch.Tensor, assign_stats: Optional[torch.Tensor],
                      mining_stats: Optional[torch.Tensor], slot: int):
        """-> stats [B_local, 4] int32; the packed shard lands in slot ``slot`` of every rank's arena."""
        n, t = int(dets.shape[0]), int(dets.shape[1])
        assert t == self.max_rows and n <= self.capacity and 0 <= slot < self.slots
        stats = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dets.device)[:n]
        N = self._N
        with torch.cuda.device(self.device):
            N.check(N.lib().ssd_pack_exchange(dets.data_ptr(), counts.data_ptr(),
                                              None if assign_stats is None else assign_stats.data_ptr(),
                                              None if mining_stats is None else mining_stats.data_ptr(), n, t,
                                              self.capacity, self._ptrs, self.world, self.rank, slot, stats.data_ptr(),
                                              self._stream()))
        return stats

    def wait(self, slot: int) -> None:
        N = self._N
        with torch.cuda.device(self.device):
            N.check(N.lib().ssd_exchange_wait(self._own, self.world, self.capacity, slot, self._stream()))

    def gathered(self, slot: int) -> torch.Tensor:
        """[world * capacity, row_words(T)] fp32 view of the slot (the layout ``unpack_gathered`` reads)."""
        words = row_words(self.max_rows)
        off = self._N.lib().ssd_exchange_slot_offset(self.world, slot, self.capacity, self.max_rows)
        n = self.world * self.capacity * words
        return self.arena[off: off + 4 * n].view(torch.float32).view(self.world * self.capacity, words)

    def unpack(self, slot: int):
        return unpack_gathered(self.gathered(slot), self.batch, self.world, self.max_rows)

    def error(self) -> int:
        """Non-zero after a peer failed to answer within the kernel's time-out (one host sync)."""
        return int(self.arena[:8].view(torch.int64)[0].item())

    def check(self) -> None:
        if self.error():
            raise RuntimeError("PeerExchange: a rank stopped answering (time-out word set by the exchange kernels)")


def unpack_gathered(gathered: torch.Tensor, batch: int, world: int, max_rows: int):
    """Inverse of pack_shard over the concatenation of all ranks' buffers."""
    capacity = gathered.shape[0] // world
    ints = gathered.view(torch.int32)
    if batch == world * capacity:                       # even shards: pure views, no kernel
        dets = gathered[:, : max_rows * 6].reshape(batch, max_rows, 6)
        return dets, ints[:, max_rows * 6], ints[:, max_rows * 6 + 1: max_rows * 6 + 5]
    keep = []
    for r in range(world):
        lo, hi = image_shard(batch, r, world)
        keep.extend(range(r * capacity, r * capacity + (hi - lo)))
    idx = torch.tensor(keep, dtype=torch.long, device=gathered.device)
    rows = gathered.index_select(0, idx)
    irows = ints.index_select(0, idx)
    dets = rows[:, : max_rows * 6].reshape(batch, max_rows, 6)
    counts = irows[:, max_rows * 6].contiguous()
    stats = irows[:, max_rows * 6 + 1: max_rows * 6 + 5].contiguous()
    return dets, counts, stats


def all_gather_detections(dets: torch.Tensor, counts: torch.Tensor, stats: torch.Tensor, batch: int,
                          group: Optional[dist.ProcessGroup] = None):
    """dets [B_local, T, 6], counts [B_local], stats [B_local, 4] -> the same for the whole batch,
    identical on every rank.  One collective."""
    world = dist.get_world_size(group)
    capacity = shard_capacity(batch, world)
    mine = pack_shard(dets, counts, stats, capacity)
    return all_gather_packed(mine, batch, dets.shape[1], group)


def all_gather_packed(mine: torch.Tensor, batch: int, max_rows: int, group: Optional[dist.ProcessGroup] = None):
    """The collective alone, for a shard already packed by :func:`pack_shard` (e.g. inside a CUDA graph)."""
    world = dist.get_world_size(group)
    capacity = mine.shape[0]
    out = torch.empty((world * capacity, mine.shape[1]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return unpack_gathered(out, batch, world, max_rows)


class OverlappedGather:
    """The exchange step off the critical path: ``submit`` enqueues the all-gather of a packed shard
    asynchronously (NCCL runs it on its own stream, after the work already enqueued on the current
    stream) and hands back the PREVIOUS submission's gathered buffer, now complete for the current
    stream -- so the collective of step i overlaps the kernels of step i+1.  ``flush`` returns the
    last one.  The buffers rotate over a small ring (no allocation per step); ``unpack`` turns one
    into ``(dets, counts, stats)`` of the whole batch as in :func:`all_gather_detections`."""

    def __init__(self, batch: int, max_rows: int, group: Optional[dist.ProcessGroup] = None, ring: int = 3):
        self.batch, self.max_rows, self.group = batch, max_rows, group
        self.world = dist.get_world_size(group)
        self._pending = None
        self._ring = [None] * ring
        self._next = 0
        self._loose = []
        self._last = None

    def unpack(self, gathered: torch.Tensor):
        return unpack_gathered(gathered, self.batch, self.world, self.max_rows)

    def _finish(self):
        if self._pending is None:
            return None
        work, out = self._pending
        self._pending = None
        work.wait()                       # orders the current stream after the collective; no host sync on NCCL
        return out

    def submit_nowait(self, shard: torch.Tensor):
        """Throughput mode (several steps in flight on several streams): enqueue the exchange behind the work
        already on the CURRENT stream and return at once; nothing is ordered after it until :meth:`flush`, which
        waits for every exchange submitted this way (NCCL runs them in submission order on its own stream, so a
        ring buffer is rewritten only after the exchange that used it before)."""
        k = self._next
        self._next = (k + 1) % len(self._ring)
        out = self._ring[k]
        shape = (self.world * shard.shape[0], shard.shape[1])
        if out is None or tuple(out.shape) != shape or out.device != shard.device:
            out = torch.empty(shape, dtype=shard.dtype, device=shard.device)
            self._ring[k] = out
        work = dist.all_gather_into_tensor(out, shard, group=self.group, async_op=True)
        self._loose.append(work)
        if len(self._loose) > 4 * len(self._ring):
            self._loose.pop(0).wait()
        self._last = out
        return out

    def submit(self, shard: torch.Tensor):
        previous = self._finish()
        k = self._next
        self._next = (k + 1) % len(self._ring)
        out = self._ring[k]
        shape = (self.world * shard.shape[0], shard.shape[1])
        if out is None or tuple(out.shape) != shape or out.device != shard.device:
            out = torch.empty(shape, dtype=shard.dtype, device=shard.device)
            self._ring[k] = out
        work = dist.all_gather_into_tensor(out, shard, group=self.group, async_op=True)
        self._pending = (work, out)
        return previous

    def flush(self):
        if self._loose:
            for work in self._loose:
                work.wait()
            self._loose = []
            return self._last
        return self._finish()
