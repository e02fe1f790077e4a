#!/usr/bin/env python
"""Target assignment alone: back-to-back launch time (CUDA events) and the device-side phase marks of one
CTA (ssd_b200_trace_enable).  Development tool: `python tools/assign_probe.py [workload]`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from single_shot_detection_b200 import _native as N  # noqa: E402
from single_shot_detection_b200 import workloads as wl  # noqa: E402
from single_shot_detection_b200.target_assigner import TargetAssigner, pack_ground_truth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else wl.HEADLINE
    w = wl.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    anchors = wl.build_anchors(w).to(dev)
    packs = []
    for s in range(4):
        _, gt, _, _ = wl.make_inputs(w, seed=23 + s)
        p = pack_ground_truth(gt, dev)
        p.rows, p.offsets = p.rows.clone(), p.offsets.clone()
        packs.append(p)
    ta = TargetAssigner(w.matched_threshold, w.unmatched_threshold, nan_check="off")
    for i in range(5):
        ta.encode_packed(packs[i % 4], anchors)
    torch.cuda.synchronize()
    iters = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ta.encode_packed(packs[i % 4], anchors)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"assign_back_to_back_us": round(1e3 * e0.elapsed_time(e1) / iters, 2)}))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(8):
            ta.encode_packed(packs[i % 4], anchors)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for i in range(25):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"assign_graph_us": round(1e3 * e0.elapsed_time(e1) / 200, 2)}))
    lib = N.lib()
    nslots = lib.ssd_b200_trace_slots()
    for r in range(3):
        buf = torch.empty((nslots, 2), dtype=torch.int64, device=dev)
        buf[:, 0] = torch.iinfo(torch.int64).max
        buf[:, 1] = 0
        torch.cuda.synchronize()
        assert lib.ssd_b200_trace_enable(buf.data_ptr()) == 0
        ta.encode_packed(packs[r], anchors)
        torch.cuda.synchronize()
        lib.ssd_b200_trace_enable(None)
        t = buf.cpu()
        t0, t1 = int(t[0, 0]), int(t[0, 1])
        marks = [round((int(t[i, 0]) - t0) / 1e3, 2) for i in range(16, nslots) if int(t[i, 1]) == 1]
        print(json.dumps({"kernel_us": round((t1 - t0) / 1e3, 2), "marks_us (staged, matched, written, ticket)": marks}))


if __name__ == "__main__":
    main()
