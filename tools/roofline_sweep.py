#!/usr/bin/env python
"""Sweep the tuning knobs of the logit-streaming kernels (SSD_TILE_BYTES, SSD_CTAS_PER_SM) for the
sampler's streaming kernel alone (back-to-back launches, rotating cold-L2 inputs), the whole sampler
and the post-processor.  Development tool: `python tools/roofline_sweep.py [workload] [iters]`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from single_shot_detection_b200 import _native as N  # noqa: E402
from single_shot_detection_b200 import workloads as wl  # noqa: E402
from single_shot_detection_b200.pipeline import AnchorPipeline  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else wl.HEADLINE
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    w = wl.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    anchors = wl.build_anchors(w)
    A, C, B = int(anchors.shape[0]), w.num_score_cols, w.batch
    per_set = B * A * (C + 4) * 4
    nsets = min(16, max(2, -(-int(1.5 * 126 * 2**20) // per_set)))
    anchors_d = anchors.to(dev)
    sets = []
    for s in range(nsets):
        _, gt, scores, locs = wl.make_inputs(w, seed=23 + s)
        sets.append((scores.to(dev), locs.to(dev), torch.zeros((B, A), dtype=torch.int64, device=dev)))
    pipe = AnchorPipeline(w.cfg())
    lib = N.lib()
    keys = torch.empty((B, A), dtype=torch.int32, device=dev)
    ws = torch.empty((max(lib.ssd_hard_negative_workspace_bytes(B, A), 256),), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    algo = B * A * (4 * C + 12)

    def timed(fn):
        for i in range(5):
            fn(i % nsets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i % nsets)
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / iters

    def keys_only(k):
        N.check(lib.ssd_mining_keys(sets[k][0].data_ptr(), sets[k][2].data_ptr(), B, A, C, keys.data_ptr(), ws.data_ptr(),
                                    ws.numel(), stream))

    from single_shot_detection_b200 import sampler

    def sampler_call(k):
        sampler.hard_negative_mining(sets[k][0].view(B, A, C), sets[k][2], 3, 5)

    def post(k):
        pipe.postprocessor.postprocess_padded((sets[k][0], sets[k][1]), anchors_d)

    for tile in (0, 6, 12, 24, 48):
        for per_sm in (0, 1, 2, 3, 4, 6, 8):
            os.environ.pop("SSD_TILE_BYTES", None)
            os.environ.pop("SSD_CTAS_PER_SM", None)
            if tile:
                os.environ["SSD_TILE_BYTES"] = str(tile * 1024)
            if per_sm:
                os.environ["SSD_CTAS_PER_SM"] = str(per_sm)
            try:
                us = timed(keys_only)
                row = {"tile_kb": tile, "ctas_per_sm": per_sm, "keys_us": round(us, 2), "GBps": round(algo / us / 1e3, 1),
                       "sampler_us": round(timed(sampler_call), 2), "post_us": round(timed(post), 2)}
            except Exception as e:  # noqa: BLE001
                row = {"tile_kb": tile, "ctas_per_sm": per_sm, "error": str(e)[:80]}
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
