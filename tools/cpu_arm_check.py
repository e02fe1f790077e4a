"""How close is the bench's CPU arm (the oracle port) to the real reference?  Build container only.

Times ``oracle.run_step`` (what ``bench.py --impl reference`` executes on the GPU box, where /root/reference
does not exist) against the reference's own modules on the same inputs and host threads:
    python tools/cpu_arm_check.py [workload] [images] [threads]
"""
import functools
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SSD_REFERENCE_ROOT", "/root/reference")
sys.modules.setdefault("jpeg4py", types.SimpleNamespace(JPEG=None))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from bf.utils import box_utils as ref_box_utils  # noqa: E402
from detection import sampler as ref_sampler  # noqa: E402
from detection.box_coder import BoxCoder  # noqa: E402
from detection.postprocessor import Postprocessor  # noqa: E402
from detection.target_assigner import TargetAssigner  # noqa: E402

from oracle import anchor_pipeline_oracle as ora  # noqa: E402
from single_shot_detection_b200 import workloads as wl  # noqa: E402


def reference_step(w, gt, anchors, scores, locs, stage):
    """The timed region of SURVEY.md 8(d) on the reference's own classes."""
    clock = time.perf_counter
    b, a = len(gt), anchors.shape[0]
    assigner = TargetAssigner(w.matched_threshold, w.unmatched_threshold)
    coder = BoxCoder(w.xy_scale, w.wh_scale, w.eps)
    post = Postprocessor(coder, w.score_threshold, {"max_per_class": w.max_per_class, "overlap_threshold": w.overlap_threshold},
                         score_converter=w.converter, max_total=w.max_total)
    t0 = clock()
    target = assigner.encode_ground_truth(gt, anchors)
    t1 = clock()
    cls = target[..., 4].long()
    logits = scores.view(b, a, -1)
    if w.sampler == "hard_negative_mining":
        ref_sampler.hard_negative_mining(logits, cls, w.ratio, w.min_neg)
    else:
        ref_sampler.naive_sampler(logits, cls)
    t2 = clock()
    tl = target[..., 0:4]
    ref_box_utils.to_centroids(tl, inplace=True)
    coder.encode_box(tl, anchors, inplace=True)
    t3 = clock()
    post.postprocess((scores, locs), anchors)
    t4 = clock()
    for key, dt in (("encode_ground_truth", t1 - t0), ("sampler", t2 - t1), ("to_centroids+encode_box", t3 - t2),
                    ("postprocess", t4 - t3)):
        stage[key] = stage.get(key, 0.0) + dt


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ssd300_voc_b32"
    images = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    w = wl.WORKLOADS[name]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=23, batch=images)
    out = {"workload": name, "images": images, "threads": threads}
    for label, fn in (("reference", lambda st: reference_step(w, gt, anchors, scores, locs, st)),
                      ("port", lambda st: ora.run_step(gt, anchors, scores, locs, w.cfg(), canonical=False,
                                                       use_torchvision=True, stage_seconds=st))):
        fn({})
        best, stage_best = None, None
        for _ in range(5):
            st = {}
            t0 = time.perf_counter()
            fn(st)
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, stage_best = dt, st
        out[label] = {"images_per_s": images / best, "stage_ms": {k: 1e3 * v for k, v in stage_best.items()}}
    out["port_over_reference"] = out["port"]["images_per_s"] / out["reference"]["images_per_s"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
