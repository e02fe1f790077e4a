#!/usr/bin/env python
"""Host -> device copy bandwidth with N ranks copying at once (run under torchrun, one rank per GPU): what the box's
PCIe / host memory can deliver to the e2e leg of bench.py, whose timed region copies 27.9 MB of pinned logits + locs per
batch and rank.  Prints per-rank and aggregate GB/s for 1, 2, 4, ... ranks copying concurrently (the others idle)."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 32 * 8732 * (21 + 4) * 4                      # logits + locs of one SSD300 b32 batch
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(4)]
    stream = torch.cuda.Stream()
    active = 1
    while active <= world:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gbs = 0.0
        if rank < active:
            iters = 200
            with torch.cuda.stream(stream):
                for i in range(8):
                    dst[i % 4].copy_(host[i % 4], non_blocking=True)
                stream.synchronize()
                t0 = time.perf_counter()
                for i in range(iters):
                    dst[i % 4].copy_(host[i % 4], non_blocking=True)
                stream.synchronize()
                dt = time.perf_counter() - t0
            gbs = nbytes * iters / dt / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        if rank == 0:
            print(f"{active} rank(s) copying: aggregate {float(t):.1f} GB/s, {float(t) / active:.1f} GB/s per rank "
                  f"(= {float(t) * 1e9 / nbytes * 32 / 1e3:.0f} k img/s of H2D alone)", flush=True)
        active *= 2
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
