#!/usr/bin/env python
"""Device-side timeline of one replayed step graph (ssd_b200_trace_enable): start / end of every
kernel relative to the first start, in microseconds.  `python tools/graph_timeline.py [workload] [reps]`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from single_shot_detection_b200 import _native as N  # noqa: E402
from single_shot_detection_b200 import workloads as wl  # noqa: E402
from single_shot_detection_b200.pipeline import AnchorPipeline  # noqa: E402
from single_shot_detection_b200.target_assigner import pack_ground_truth  # noqa: E402

NAMES = {0: "assign", 1: "mining_loss", 2: "mining_keys", 3: "mining_select", 4: "pass1", 5: "gate", 6: "pass2",
         7: "nms", 8: "topk", 9: "misc"}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else wl.HEADLINE
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    w = wl.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    anchors = wl.build_anchors(w).to(dev)
    A, C, B = int(anchors.shape[0]), w.num_score_cols, w.batch
    nsets = min(16, max(2, -(-int(1.5 * 126 * 2**20) // (B * A * (C + 4) * 4))))
    pipes = []
    for s in range(nsets):
        _, gt, scores, locs = wl.make_inputs(w, seed=23 + s)
        packed = pack_ground_truth(gt, dev)
        packed.rows, packed.offsets = packed.rows.clone(), packed.offsets.clone()
        p = AnchorPipeline(w.cfg())
        p.inputs = (packed, anchors, scores.to(dev), locs.to(dev))     # the graph holds raw pointers: keep them alive
        p.capture(*p.inputs)
        pipes.append(p)
    lib = N.lib()
    nslots = lib.ssd_b200_trace_slots()
    for r in range(3 * nsets):
        pipes[r % nsets].replay()
    torch.cuda.synchronize()
    for r in range(reps):
        buf = torch.empty((nslots, 2), dtype=torch.int64, device=dev)
        buf[:, 0] = torch.iinfo(torch.int64).max
        buf[:, 1] = 0
        for q in range(nsets - 1):                      # evict the set about to run from L2
            pipes[(r + 1 + q) % nsets].replay()
        torch.cuda.synchronize()
        assert lib.ssd_b200_trace_enable(buf.data_ptr()) == 0
        pipes[r % nsets].replay()
        torch.cuda.synchronize()
        lib.ssd_b200_trace_enable(None)
        t = buf.cpu()
        used = [(int(t[i, 0]), int(t[i, 1]), i) for i in range(min(nslots, 16)) if int(t[i, 1]) > 0]
        marks = [int(t[i, 0]) for i in range(16, nslots) if int(t[i, 1]) == 1]
        t0 = min(u[0] for u in used)
        line = {NAMES.get(i, f"box_op{i - 10}"): [round((a - t0) / 1e3, 1), round((b - t0) / 1e3, 1)] for a, b, i in sorted(used)}
        line["span_us"] = round((max(u[1] for u in used) - t0) / 1e3, 1)
        if marks:
            line["marks_us"] = [round((m - t0) / 1e3, 1) for m in marks]
        print(json.dumps(line))


if __name__ == "__main__":
    main()
