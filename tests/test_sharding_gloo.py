"""Multi-rank logic on the CPU: image sharding and the one exchange step over the
gloo backend, world_size 2 and 3.  Invariant: the gathered result equals the unsharded result
bit for bit, on every rank, for even and ragged shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from single_shot_detection_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _whole_batch(batch, max_rows, seed=5):
    g = torch.Generator().manual_seed(seed)
    dets = torch.randn((batch, max_rows, 6), generator=g)
    counts = torch.randint(0, max_rows + 1, (batch,), generator=g, dtype=torch.int32)
    stats = torch.randint(0, 9000, (batch, 4), generator=g, dtype=torch.int32)
    return dets, counts, stats


def _worker(rank, world, port, batch, max_rows, failures):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dets, counts, stats = _whole_batch(batch, max_rows)
        lo, hi = sharding.image_shard(batch, rank, world)
        got = sharding.all_gather_detections(dets[lo:hi].clone(), counts[lo:hi].clone(), stats[lo:hi].clone(), batch)
        ok = (torch.equal(got[0], dets) and torch.equal(got[1].to(torch.int32), counts)
              and torch.equal(got[2].to(torch.int32), stats))
        if not ok:
            failures.put((rank, "gathered result differs from the unsharded one"))
        # the pre-packed variant used inside CUDA graphs gives the same answer
        cap = sharding.shard_capacity(batch, world)
        mine = sharding.pack_shard(dets[lo:hi], counts[lo:hi], stats[lo:hi], cap)
        got2 = sharding.all_gather_packed(mine, batch, max_rows)
        if not all(torch.equal(a.to(b.dtype), b) for a, b in zip(got2, (dets, counts, stats))):
            failures.put((rank, "all_gather_packed differs"))
        # the overlapped form hands back the previous submission's result, flush the last one
        og = sharding.OverlappedGather(batch, max_rows)
        dets_b, counts_b, stats_b = _whole_batch(batch, max_rows, seed=6)
        mine_b = sharding.pack_shard(dets_b[lo:hi], counts_b[lo:hi], stats_b[lo:hi], cap)
        first = og.submit(mine)
        second = og.submit(mine_b)
        third = og.flush()
        if first is not None or og.flush() is not None:
            failures.put((rank, "OverlappedGather: unexpected result before / after the queue"))
        for got3, want in ((second, (dets, counts, stats)), (third, (dets_b, counts_b, stats_b))):
            if not all(torch.equal(a.to(b.dtype), b) for a, b in zip(og.unpack(got3), want)):
                failures.put((rank, "OverlappedGather differs"))
        # throughput mode: several exchanges enqueued without waiting, one flush at the end
        og2 = sharding.OverlappedGather(batch, max_rows, ring=4)
        bufs = [og2.submit_nowait(m) for m in (mine, mine_b, mine)]
        last = og2.flush()
        if last is not bufs[-1]:
            failures.put((rank, "submit_nowait: flush must return the last buffer"))
        for got4, want in zip(bufs, ((dets, counts, stats), (dets_b, counts_b, stats_b), (dets, counts, stats))):
            if not all(torch.equal(a.to(b.dtype), b) for a, b in zip(og2.unpack(got4), want)):
                failures.put((rank, "submit_nowait differs"))
    except Exception as e:  # noqa: BLE001
        failures.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 8), (2, 5), (3, 7), (2, 1)])
def test_all_gather_equals_unsharded(world, batch):
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, 7, failures)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    assert failures.empty(), failures.get()


def test_image_shard_partitions_the_batch():
    for batch in (0, 1, 5, 32, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.image_shard(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            assert max(sizes) <= sharding.shard_capacity(batch, world) or batch == 0


def test_pack_unpack_roundtrip_single_rank():
    dets, counts, stats = _whole_batch(6, 5)
    buf = sharding.pack_shard(dets, counts, stats, 6)
    d, c, s = sharding.unpack_gathered(buf, 6, 1, 5)
    assert torch.equal(d, dets) and torch.equal(c, counts) and torch.equal(s, stats)
    # padding rows of a short shard are marked with count -1
    short = sharding.pack_shard(dets[:2], counts[:2], stats[:2], 4)
    assert short.view(torch.int32)[2:, 5 * 6].tolist() == [-1, -1]
