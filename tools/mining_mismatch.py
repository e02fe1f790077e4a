"""From-logits hard-negative mining against the reference's CPU criterion, per BASELINE config.

The selection kernel is bit exact on identical fp32 losses (tests/test_gpu_parity.py); from the LOGITS the
criterion -log_softmax(x)[0] is computed on the device with MUFU ex2 / lg2 (csrc/rowstream.cuh), a few ulp away
from ATen's CPU log_softmax, so an anchor whose loss sits within those ulps of the cut can fall on the other side.
This tool counts such anchors: `python tools/mining_mismatch.py [batch]` (needs a GPU; the oracle runs on the host).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import anchor_pipeline_oracle as ora  # noqa: E402
from single_shot_detection_b200 import sampler, workloads as wl  # noqa: E402
from single_shot_detection_b200.pipeline import AnchorPipeline  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda", 0)
    out = {}
    for name in ["ssd300_voc_b32", "ssd_mb2_coco_b64", "ssd512_coco_b32", "m2det512_coco_b256"]:
        w = wl.WORKLOADS[name]
        rows = []
        for seed in (23, 24, 25, 26):
            anchors, gt, scores, locs = wl.make_inputs(w, seed=seed, batch=batch)
            a, c = anchors.shape[0], w.num_score_cols
            target = ora.assign_targets(gt, anchors, w.matched_threshold, w.unmatched_threshold)
            cls = target[..., 4].long()
            logits = scores.view(batch, a, c)
            ref = ora.mine_hard_negatives(logits, cls, w.ratio, w.min_neg, canonical=True)
            mask = sampler.hard_negative_mining(logits.to(dev), cls.to(dev), w.ratio, w.min_neg).cpu()
            # the eval step takes the criterion from the post-processor's first pass instead
            pipe = AnchorPipeline(w.cfg())
            _, mask2, _ = pipe.step(gt, anchors, scores, locs)
            rows.append((int((mask != ref).sum()), int((mask2.cpu() != ref).sum()), int(ref.sum())))
        out[name] = {"images": 4 * batch, "anchors_selected": sum(r[2] for r in rows),
                     "mismatch_sampler_kernel": sum(r[0] for r in rows),
                     "mismatch_shared_pass": sum(r[1] for r in rows)}
        print(name, out[name], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
