#!/usr/bin/env python
"""Per-stage device times of the anchor pipeline (CUDA events, inputs rotating over sets larger
than L2).  Development tool: `python tools/stage_times.py [workload] [iters]` on the GPU box.

Each stage is one reference-shaped API call (one or more launches); the last lines are the whole
step, eager and replayed from a CUDA graph.  Prints one JSON object.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from single_shot_detection_b200 import _native as N  # noqa: E402
from single_shot_detection_b200 import box_utils, workloads as wl  # noqa: E402
from single_shot_detection_b200.pipeline import AnchorPipeline  # noqa: E402
from single_shot_detection_b200.target_assigner import pack_ground_truth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else wl.HEADLINE
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
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
        packed = pack_ground_truth(gt, dev)
        packed.rows, packed.offsets = packed.rows.clone(), packed.offsets.clone()
        sets.append((packed, scores.to(dev), locs.to(dev)))
    torch.cuda.synchronize()
    pipe = AnchorPipeline(w.cfg())
    targets = [pipe.target_assigner.encode_packed(p, anchors_d) for p, _, _ in sets]
    classes = [t[..., 4].long() for t in targets]
    torch.cuda.synchronize()

    def timed(fn):
        for i in range(3):
            fn(i % nsets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = N.lib().ssd_b200_launch_count()
        e0.record()
        for i in range(iters):
            fn(i % nsets)
        e1.record()
        torch.cuda.synchronize()
        launches = (N.lib().ssd_b200_launch_count() - l0) / iters
        return {"us": round(1e3 * e0.elapsed_time(e1) / iters, 2), "launches": launches}

    out = {"workload": name, "A": A, "C": C, "B": B, "nsets": nsets}
    out["assign"] = timed(lambda k: pipe.target_assigner.encode_packed(sets[k][0], anchors_d))
    out["sampler"] = timed(lambda k: pipe.sampler(sets[k][1].view(B, A, C), classes[k]))
    out["to_centroids+encode"] = timed(lambda k: (box_utils.to_centroids(targets[k][..., :4], inplace=True),
                                                  pipe.box_coder.encode_box(targets[k][..., :4], anchors_d, inplace=True)))
    out["postprocess"] = timed(lambda k: pipe.postprocessor.postprocess_padded((sets[k][1], sets[k][2]), anchors_d))
    out["step_eager"] = timed(lambda k: pipe.step_device(sets[k][0], anchors_d, sets[k][1], sets[k][2]))
    pipes = [AnchorPipeline(w.cfg()) for _ in sets]
    for p_, (pk, sc, lc) in zip(pipes, sets):
        p_.capture(pk, anchors_d, sc, lc)
    out["step_graph"] = timed(lambda k: pipes[k].replay())
    out["img_per_s_graph"] = round(B / (out["step_graph"]["us"] * 1e-6))
    # per-kernel times in context (library event timers, eager step, warm-up discarded)
    import ctypes
    lib = N.lib()
    for i in range(3):
        pipe.step_device(sets[i % nsets][0], anchors_d, sets[i % nsets][1], sets[i % nsets][2])
    torch.cuda.synchronize()
    lib.ssd_b200_timing_enable(1)
    for i in range(20):
        k = i % nsets
        pipe.step_device(sets[k][0], anchors_d, sets[k][1], sets[k][2])
    buf = ctypes.create_string_buffer(4096)
    lib.ssd_b200_timing_report(buf, 4096)
    lib.ssd_b200_timing_enable(0)
    out["kernels_us"] = {kv.split(":")[0]: float(kv.split(":")[1]) for kv in buf.value.decode().split(",") if kv}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
