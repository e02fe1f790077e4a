#!/usr/bin/env python
"""Multi-GPU check of the peer-memory exchange (run under torchrun, one rank per GPU):
`python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_check.py`
Every rank packs random shards into several slots for several rounds with sharding.PeerExchange and compares each
gathered slot, bit for bit, with NCCL's all-gather of the tensor-op packing (sharding.all_gather_detections).
A second phase skews the ranks: one rank per round stalls its stream between wait() and its read of the gathered
slots while the others run ahead into the next rounds -- the slot must still hold the round it waited for (the
writers wait for every rank's open() of the next launch before they overwrite)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from single_shot_detection_b200 import sharding  # noqa: E402
from single_shot_detection_b200.pipeline import matched_stats  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # XCHK_BATCH / XCHK_T: the shard shape (default a small one; 32 / 200 is the headline step's, used for the ncu capture)
    batch_local, T, slots = int(os.environ.get("XCHK_BATCH", "5")), int(os.environ.get("XCHK_T", "40")), 3
    batch = batch_local * world
    px = sharding.PeerExchange(batch, T, slots=slots)
    gen = torch.Generator().manual_seed(100 + rank)
    ok = True
    for rnd in range(6):
        want = []
        for k in range(slots):
            dets = torch.rand((batch_local, T, 6), generator=gen).to(dev)
            counts = torch.randint(0, T + 1, (batch_local,), generator=gen, dtype=torch.int32).to(dev)
            a_stats = torch.randint(0, 99, (batch_local, 4), generator=gen, dtype=torch.int32).to(dev)
            m_stats = torch.randint(0, 99, (batch_local, 4), generator=gen, dtype=torch.int32).to(dev)
            px.open(k)
            px.pack_exchange(dets, counts, a_stats, m_stats, k)
            want.append(sharding.all_gather_detections(dets, counts, matched_stats(a_stats, m_stats, counts), batch))
        for k in range(slots):
            px.wait(k)
        torch.cuda.synchronize()
        for k in range(slots):
            got = px.unpack(k)
            for g, r in zip(got, want[k]):
                if not torch.equal(g.to(r.dtype), r):
                    ok = False
                    print(f"rank {rank} round {rnd} slot {k}: gathered buffer differs", flush=True)
        dist.barrier()
    # ---- skewed ranks: the reader of round r is slow, the writers of round r + 1 must wait for it ----
    snaps, wants = [], []
    for rnd in range(8):
        want = []
        for k in range(slots):
            dets = torch.rand((batch_local, T, 6), generator=gen).to(dev)
            counts = torch.randint(0, T + 1, (batch_local,), generator=gen, dtype=torch.int32).to(dev)
            a_stats = torch.randint(0, 99, (batch_local, 4), generator=gen, dtype=torch.int32).to(dev)
            px.open(k)
            px.pack_exchange(dets, counts, a_stats, None, k)
            want.append(sharding.all_gather_detections(dets, counts, matched_stats(a_stats, None, counts), batch))
        for k in range(slots):
            px.wait(k)
        if rnd % world == rank:
            torch.cuda._sleep(40_000_000)                  # ~20 ms: the peers are rounds ahead by now
        snaps.append([px.gathered(k).clone() for k in range(slots)])
        wants.append(want)
    torch.cuda.synchronize()
    for rnd, (snap, want) in enumerate(zip(snaps, wants)):
        for k in range(slots):
            got = sharding.unpack_gathered(snap[k], batch, world, T)
            for g, r in zip(got, want[k]):
                if not torch.equal(g.to(r.dtype), r):
                    ok = False
                    print(f"rank {rank} skewed round {rnd} slot {k}: slot overwritten while it was being read", flush=True)
    dist.barrier()
    err = px.error()
    px.close()
    flag = torch.tensor([0 if ok and err == 0 else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("exchange_check", "OK" if int(flag) == 0 else "FAILED", f"(world {world}, error word {err})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()
