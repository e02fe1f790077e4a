"""Generate the golden fixtures in this directory by running the REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own hot-path modules (``detection.target_assigner``,
``detection.matcher``, ``detection.box_coder``, ``detection.sampler``,
``detection.postprocessor``, ``bf.utils.box_utils``, ``detection.anchor_generators``,
``detection.losses.multibox_loss``) with a dummy ``jpeg4py`` module (needed only because
``bf/datasets/detection_dataset.py:3`` imports it), feeds them seeded synthetic inputs and
stores inputs + outputs as ``*.npz``.  The fixtures pin ``oracle/anchor_pipeline_oracle.py``
(tests/test_oracle_golden.py) and are compared directly with the CUDA path (tests/test_gpu_*.py).

Versions at generation time are recorded in ``golden_meta.json``.
"""
from __future__ import annotations

import functools
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SSD_REFERENCE_ROOT", "/root/reference")

sys.modules.setdefault("jpeg4py", types.SimpleNamespace(JPEG=None))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import torchvision  # noqa: E402
from bf.utils import box_utils as ref_box_utils  # noqa: E402
from detection import matcher as ref_matcher  # noqa: E402
from detection import sampler as ref_sampler  # noqa: E402
from detection.anchor_generators import retina_net as ref_retina  # noqa: E402
from detection.anchor_generators import ssd as ref_ssd  # noqa: E402
from detection.box_coder import BoxCoder as RefBoxCoder  # noqa: E402
from detection.losses.multibox_loss import MultiboxLoss as RefMultiboxLoss  # noqa: E402
from detection.postprocessor import Postprocessor as RefPostprocessor  # noqa: E402
from detection.target_assigner import TargetAssigner as RefTargetAssigner  # noqa: E402

from single_shot_detection_b200 import workloads as wl  # noqa: E402


def ragged(list_of_tensors):
    """list of [n_i, k] -> (flat [sum n_i, k], offsets [len+1])."""
    k = max((t.shape[1] for t in list_of_tensors), default=6)
    sizes = [t.shape[0] for t in list_of_tensors]
    flat = torch.cat([t.reshape(-1, k) for t in list_of_tensors], dim=0) if sizes else torch.zeros((0, k))
    return flat.numpy(), np.cumsum([0] + sizes).astype(np.int64)


def reference_anchors(w: wl.Workload) -> torch.Tensor:
    class _Shape:                       # stands in for a [B, C, H, W] tensor
        def __init__(self, h, wd):
            self._s = (1, 1, h, wd)

        def size(self, i):
            return self._s[i]

    if w.anchor_kind == "ssd":
        lo, hi, ratios = w.anchor_args
        gens = ref_ssd.build_anchor_generators(num_scales=len(w.fmaps), min_scale=lo, max_scale=hi,
                                               aspect_ratios=[list(r) for r in ratios])
    else:
        ratios, min_level, scale, spl = w.anchor_args
        gens = ref_retina.build_anchor_generators(aspect_ratios=list(ratios), min_level=min_level,
                                                  max_level=min_level + len(w.fmaps) - 1,
                                                  scale=scale, scales_per_level=spl)
    img = _Shape(w.img, w.img)
    parts = [g.generate(img, _Shape(s, s)).view(-1) for g, s in zip(gens, w.fmaps)]
    return torch.cat(parts, dim=0).view(-1, 4)


def run_pipeline_case(name: str, w: wl.Workload, gt, anchors, scores, locs, out_dir: str, full: bool = True):
    """Run every reference stage on one set of inputs and save inputs + stage outputs."""
    b, a = len(gt), anchors.shape[0]
    c = w.num_score_cols
    assigner = RefTargetAssigner(w.matched_threshold, w.unmatched_threshold)
    target = assigner.encode_ground_truth(gt, anchors)

    corner_anchors = ref_box_utils.to_corners(anchors)
    match_idx = torch.full((b, a), ref_matcher.NOT_MATCHED, dtype=torch.long)
    iou0 = np.zeros((0, a), dtype=np.float32)
    for i, g in enumerate(gt):
        if len(g):
            iou = ref_box_utils.iou(g[:, :4], corner_anchors)
            match_idx[i] = ref_matcher.match_per_prediction(iou, w.matched_threshold, w.unmatched_threshold)
            if i == 0:
                iou0 = iou.numpy()

    cls = target[..., 4].long()
    logits = scores.view(b, a, c)
    neg_loss = -torch.nn.functional.log_softmax(logits, dim=-1)[:, :, 0]
    hnm_mask = ref_sampler.hard_negative_mining(logits, cls, w.ratio, w.min_neg)
    naive_mask = ref_sampler.naive_sampler(logits, cls)

    coder = RefBoxCoder(w.xy_scale, w.wh_scale, w.eps)
    enc_inplace = target.clone()
    tl = enc_inplace[..., 0:4]
    ref_box_utils.to_centroids(tl, inplace=True)
    centroids_inplace = tl.clone()
    coder.encode_box(tl, anchors, inplace=True)
    centroids_oop = ref_box_utils.to_centroids(target[..., 0:4])
    enc_oop = coder.encode_box(centroids_oop, anchors)
    decoded = coder.decode_box(locs.view(b, a, 4), anchors, inplace=torch.tensor(0))
    decoded_inplace = coder.decode_box(locs.view(b, a, 4).clone(), anchors, inplace=torch.tensor(1))
    decoded_corners = ref_box_utils.to_corners(decoded)

    if w.converter == "SOFTMAX":
        probs = torch.nn.functional.softmax(logits, dim=-1)
    else:
        probs = torch.sigmoid(logits)

    post = RefPostprocessor(coder, w.score_threshold,
                            {"max_per_class": w.max_per_class, "overlap_threshold": w.overlap_threshold},
                            score_converter=w.converter, max_total=w.max_total)
    dets = post.postprocess((scores, locs), anchors)
    post_all = RefPostprocessor(coder, w.score_threshold,
                                {"max_per_class": w.max_per_class, "overlap_threshold": w.overlap_threshold},
                                score_converter=w.converter, max_total=None)
    dets_all = post_all.postprocess((scores, locs), anchors)

    # the caller's loss with the reference sampler / coder plugged in (multibox_loss.py:35-94)
    if w.converter == "SOFTMAX":
        sampler = functools.partial(ref_sampler.hard_negative_mining,
                                    negative_per_positive_ratio=w.ratio, min_negative_per_image=w.min_neg)
        crit = RefMultiboxLoss(sampler, coder, {"name": "CrossEntropyLoss"}, {"name": "SmoothL1Loss"})
    else:
        crit = RefMultiboxLoss(ref_sampler.naive_sampler, coder,
                               {"name": "SigmoidFocalLoss", "gamma": 2.0, "alpha": 0.25},
                               {"name": "SmoothL1Loss"})
    loss3 = [float(x) for x in crit((scores, locs), anchors, target.clone())]

    gt_flat, gt_off = ragged(gt)
    det_flat, det_off = ragged(dets)
    det_all_flat, det_all_off = ragged(dets_all)
    blob = dict(
        workload=np.array(w.name), anchors=anchors.numpy(), gt_flat=gt_flat, gt_off=gt_off,
        gt_cols=np.array(gt[0].shape[1] if gt else 6),
        thresholds=np.array([w.matched_threshold, w.unmatched_threshold], dtype=np.float64),
        scores=scores.numpy(), locs=locs.numpy(),
        target=target.numpy(), match_idx=match_idx.numpy(),
        hnm_mask=np.packbits(hnm_mask.numpy(), axis=1), naive_mask=np.packbits(naive_mask.numpy(), axis=1),
        enc_inplace=enc_inplace[..., 0:4].numpy(),
        det_flat=det_flat, det_off=det_off, det_all_flat=det_all_flat, det_all_off=det_all_off,
        loss3=np.array(loss3, dtype=np.float64))
    if full:        # float-heavy intermediates only for the small cases (fixture size)
        blob.update(
            iou0=iou0, neg_loss=neg_loss.numpy(),
            centroids_inplace=centroids_inplace.numpy(), centroids_oop=centroids_oop.numpy(),
            enc_oop=enc_oop.numpy(), decoded=decoded.numpy(), decoded_inplace=decoded_inplace.numpy(),
            decoded_corners=decoded_corners.numpy(), probs=probs.numpy())
    np.savez_compressed(os.path.join(out_dir, f"pipeline_{name}.npz"), **blob)
    print(f"  pipeline_{name}: B={b} A={a} C={c} dets={[int(d.shape[0]) for d in dets]}")


# BASELINE.json configs at their full anchor / class shapes (B = 2 images each).  The inputs are NOT stored --
# random head outputs do not compress (16 MB for SSD512) -- they are regenerated by
# ``workloads.make_inputs(w, seed, batch)`` and pinned by SHA-256; only the reference's compact outputs
# are kept: matched indices (int16), targets (mostly zero rows), bit-packed masks, detections, the anchor of
# every kept row, the coded boxes of the matched rows + a sample of the others, the loss triple.
CONFIG_CASES = [("ssd300_voc_b32", 2, 23), ("ssd512_coco_b32", 2, 23), ("retina500_coco_b32", 2, 23),
                ("m2det512_coco_b256", 2, 23)]


def _sha(t: torch.Tensor) -> str:
    import hashlib
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def run_config_case(name: str, batch: int, seed: int, out_dir: str):
    w = wl.WORKLOADS[name]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=seed, batch=batch)
    ref_anchors = reference_anchors(w)
    assert torch.equal(ref_anchors, anchors), name          # the package's table IS the reference's (anchors.npz)
    b, a, c = batch, anchors.shape[0], w.num_score_cols
    assigner = RefTargetAssigner(w.matched_threshold, w.unmatched_threshold)
    target = assigner.encode_ground_truth(gt, anchors)
    corner_anchors = ref_box_utils.to_corners(anchors)
    match_idx = torch.full((b, a), ref_matcher.NOT_MATCHED, dtype=torch.long)
    for i, g in enumerate(gt):
        if len(g):
            match_idx[i] = ref_matcher.match_per_prediction(ref_box_utils.iou(g[:, :4], corner_anchors),
                                                            w.matched_threshold, w.unmatched_threshold)
    cls = target[..., 4].long()
    logits = scores.view(b, a, c)
    # detection/init.py:90-92 builds the sampler from the config; the reference function is called directly here
    hnm_mask = ref_sampler.hard_negative_mining(logits, cls, w.ratio, w.min_neg)
    naive_mask = ref_sampler.naive_sampler(logits, cls)

    coder = RefBoxCoder(w.xy_scale, w.wh_scale, w.eps)
    enc = target.clone()
    tl = enc[..., 0:4]
    ref_box_utils.to_centroids(tl, inplace=True)                       # multibox_loss.py:81
    coder.encode_box(tl, anchors, inplace=True)                        # multibox_loss.py:82
    flat_cls = target[..., 4].reshape(-1)
    rows = torch.nonzero(flat_cls != 0).flatten()
    sample = torch.arange(0, b * a, 37)
    enc_rows = torch.unique(torch.cat([rows, sample]))
    enc_vals = tl.reshape(-1, 4)[enc_rows]

    nms_cfg = {"max_per_class": w.max_per_class, "overlap_threshold": w.overlap_threshold}
    post = RefPostprocessor(coder, w.score_threshold, nms_cfg, score_converter=w.converter, max_total=w.max_total)
    dets = post.postprocess((scores, locs), anchors)
    post_all = RefPostprocessor(coder, w.score_threshold, nms_cfg, score_converter=w.converter, max_total=None)
    dets_all = post_all.postprocess((scores, locs), anchors)

    # the anchor behind every kept row, from the reference's own intermediates (postprocessor.py:44-55)
    probs = post.score_converter_fn(logits)
    fg = probs[..., 1:] if w.converter == "SOFTMAX" else probs
    corners = ref_box_utils.to_corners(coder.decode_box(locs.view(b, a, 4), anchors, inplace=torch.tensor(0)))
    det_anchor = []
    for i, d in enumerate(dets_all):
        out = torch.empty((d.shape[0],), dtype=torch.int32)
        for r in range(d.shape[0]):
            col = int(d[r, 4]) - 1
            hit = torch.nonzero((fg[i, :, col] == d[r, 5]) & (corners[i] == d[r, :4]).all(dim=1)).flatten()
            assert hit.numel() >= 1, (name, i, r)
            out[r] = int(hit[0])
        det_anchor.append(out)

    if w.converter == "SOFTMAX":
        sampler = functools.partial(ref_sampler.hard_negative_mining,
                                    negative_per_positive_ratio=w.ratio, min_negative_per_image=w.min_neg)
        crit = RefMultiboxLoss(sampler, coder, {"name": "CrossEntropyLoss"}, {"name": "SmoothL1Loss"})
    else:
        crit = RefMultiboxLoss(ref_sampler.naive_sampler, coder,
                               {"name": "SigmoidFocalLoss", "gamma": 2.0, "alpha": 0.25}, {"name": "SmoothL1Loss"})
    loss3 = [float(x) for x in crit((scores, locs), anchors, target.clone())]

    gt_flat, gt_off = ragged(gt)
    det_flat, det_off = ragged(dets)
    det_all_flat, det_all_off = ragged(dets_all)
    assert int(match_idx.max()) < 32767
    blob = dict(
        workload=np.array(w.name), regen_seed=np.array(seed), regen_batch=np.array(batch),
        sha_scores=np.array(_sha(scores)), sha_locs=np.array(_sha(locs)), sha_anchors=np.array(_sha(anchors)),
        gt_flat=gt_flat, gt_off=gt_off, gt_cols=np.array(6),
        thresholds=np.array([w.matched_threshold, w.unmatched_threshold], dtype=np.float64),
        target=target.numpy(), match_idx=match_idx.numpy().astype(np.int16),
        hnm_mask=np.packbits(hnm_mask.numpy(), axis=1), naive_mask=np.packbits(naive_mask.numpy(), axis=1),
        enc_rows=enc_rows.numpy().astype(np.int32), enc_vals=enc_vals.numpy(),
        det_flat=det_flat, det_off=det_off, det_all_flat=det_all_flat, det_all_off=det_all_off,
        det_all_anchor=torch.cat(det_anchor).numpy(), loss3=np.array(loss3, dtype=np.float64))
    np.savez_compressed(os.path.join(out_dir, f"pipeline_{name}.npz"), **blob)
    print(f"  pipeline_{name}: B={b} A={a} C={c} dets={[int(d.shape[0]) for d in dets]} "
          f"kept={[int(d.shape[0]) for d in dets_all]} positives={int((hnm_mask & naive_mask).sum())}")


def edge_case_inputs(w: wl.Workload, anchors: torch.Tensor, gen: torch.Generator):
    """Hand-made ground truth exercising the matcher's tie rules and the API edge cases."""
    img = float(w.img)
    a0 = ref_box_utils.to_corners(anchors[:1])[0]            # exactly anchor 0
    mid = ref_box_utils.to_corners(anchors[anchors.shape[0] // 2: anchors.shape[0] // 2 + 1])[0]
    gt = []
    # image 0: empty GT (target_assigner.py:43-44)
    gt.append(torch.zeros((0, 6)))
    # image 1: duplicated GT rows (mixup, fractional scores) -> per-anchor tie -> lowest GT index;
    #          forced-match collision -> highest GT index
    rows = torch.tensor([[*mid.tolist(), 2.0, 0.7], [*mid.tolist(), 3.0, 0.3],
                         [*a0.tolist(), 1.0, 1.0]])
    gt.append(rows)
    # image 2: a GT entirely outside every anchor's reach (all-zero IoU row -> anchor 0 forced)
    #          plus a normal one, 7 columns (difficult flag) as voc.py:45-53 emits
    far = torch.tensor([[img * 50, img * 50, img * 50 + 3, img * 50 + 3, 4.0, 1.0, 1.0],
                        [img * .2, img * .2, img * .7, img * .6, 5.0, 1.0, 0.0]])
    gt.append(far)
    # image 3: many overlapping GTs so several force the same anchors
    base = wl.make_ground_truth(1, w.img, w.num_fg, w.max_gt, gen)[0]
    jitter = base.clone()
    jitter[:, :4] += 0.25
    gt.append(torch.cat([base, jitter, base], dim=0))
    return gt


def nms_cases(gen: torch.Generator):
    """torchvision.ops.nms known answers (the third-party arithmetic of box_utils.py:193)."""
    cases = []
    # the two probes quoted in SURVEY.md §8(c)
    cases.append((torch.tensor([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [0, 0, 10, 9]], dtype=torch.float32),
                  torch.tensor([.8, .8, .9, .8]), .45))
    cases.append((torch.tensor([[0, 0, 10, 10], [0, 0, 4, 10]], dtype=torch.float32),
                  torch.tensor([.9, .8]), 0.4))
    cases.append((torch.zeros((0, 4)), torch.zeros((0,)), 0.5))
    for n, thr in [(1, .5), (7, .3), (100, .45), (100, .5), (257, .45), (64, .0), (64, 1.0), (300, .6)]:
        c = torch.rand((n, 2), generator=gen) * 100
        s = torch.rand((n, 2), generator=gen) * 40 + 1
        boxes = torch.cat([c - s / 2, c + s / 2], dim=1)
        scores = torch.rand((n,), generator=gen)
        if n >= 64:                              # exact score ties and duplicate boxes
            scores[n // 2:] = scores[: n - n // 2].clone()
            boxes[n // 4: n // 4 + 8] = boxes[:8]
        cases.append((boxes, scores, thr))
    # degenerate: inverted boxes (negative unclamped area) and zero-area boxes
    boxes = torch.tensor([[10, 10, 5, 5], [0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 8, 8], [2, 2, 9, 9]], dtype=torch.float32)
    cases.append((boxes, torch.tensor([.5, .9, .8, .7, .6]), .3))
    blob = {"num_cases": np.array(len(cases))}
    for i, (bx, sc, thr) in enumerate(cases):
        keep = torchvision.ops.nms(bx, sc, thr)
        blob[f"boxes_{i}"] = bx.numpy()
        blob[f"scores_{i}"] = sc.numpy()
        blob[f"thr_{i}"] = np.array(thr, dtype=np.float64)
        blob[f"keep_{i}"] = keep.numpy()
    # box_utils.nms with top-k (bf/utils/box_utils.py:165-194)
    for i, (n, k) in enumerate([(500, 100), (80, 100), (101, 100)]):
        c = torch.rand((n, 2), generator=gen) * 60
        s = torch.rand((n, 2), generator=gen) * 30 + 1
        bx = torch.cat([c - s / 2, c + s / 2], dim=1)
        sc = torch.rand((n,), generator=gen)
        (bk, sk), keep = ref_box_utils.nms(bx, sc, overlap_threshold=.45, score_threshold=.01, max_per_class=k)
        blob[f"topk_boxes_{i}"] = bx.numpy()
        blob[f"topk_scores_{i}"] = sc.numpy()
        blob[f"topk_k_{i}"] = np.array(k)
        blob[f"topk_kept_boxes_{i}"] = bk.numpy()
        blob[f"topk_kept_scores_{i}"] = sk.numpy()
    blob["num_topk_cases"] = np.array(3)
    return blob


def main():
    out_dir = HERE
    torch.manual_seed(23)
    torch.set_num_threads(1)
    if "--configs-only" in sys.argv:
        print("BASELINE config cases")
        for name, batch, seed in CONFIG_CASES:
            run_config_case(name, batch, seed, out_dir)
        return

    print("anchor tables")
    blob = {}
    for name in ["ssd300_voc_b8", "ssd300_voc_8108_b8", "ssd_mb2_coco_b64", "ssd512_coco_b32",
                 "retina500_coco_b32", "m2det512_coco_b256", "tiny_voc_b3", "tiny_sigmoid_b2"]:
        w = wl.WORKLOADS[name]
        ref = reference_anchors(w)
        mine = wl.build_anchors(w)
        assert ref.shape == mine.shape, (name, ref.shape, mine.shape)
        blob[name] = ref.numpy()
        print(f"  {name}: A={ref.shape[0]} bit-equal to package builder: {bool(torch.equal(ref, mine))}")
    np.savez_compressed(os.path.join(out_dir, "anchors.npz"), **blob)

    print("pipeline cases")
    gen = torch.Generator().manual_seed(23)
    for name, batch in [("tiny_voc_b3", None), ("tiny_sigmoid_b2", None),
                        ("ssd_mb2_coco_b64", 2), ("ssd300_voc_b8", 2)]:
        w = wl.WORKLOADS[name]
        anchors, gt, scores, locs = wl.make_inputs(w, seed=23, batch=batch)
        anchors = reference_anchors(w)
        run_pipeline_case(name, w, gt, anchors, scores, locs, out_dir, full=anchors.shape[0] < 1000)

    # edge cases on the tiny workload
    w = wl.WORKLOADS["tiny_voc_b3"]
    anchors = reference_anchors(w)
    gt = edge_case_inputs(w, anchors, gen)
    scores, locs = wl.make_head_outputs(len(gt), anchors.shape[0], w.num_score_cols, 0.0, gen)
    # 7-column rows and 6-column rows cannot share one ragged array: pad 6 -> 7 columns
    gt7 = [torch.cat([g, torch.zeros((g.shape[0], 7 - g.shape[1]))], dim=1) for g in gt]
    run_pipeline_case("edge_tiny", w, gt7, anchors, scores, locs, out_dir)

    # mixup ground truth + different thresholds (ignore band) on the tiny workload
    w2 = wl.Workload(**{**w.__dict__, "name": "tiny_voc_b3", "matched_threshold": 0.5, "unmatched_threshold": 0.3})
    gt = wl.make_ground_truth(3, w2.img, w2.num_fg, w2.max_gt, gen, mixup=0.3)
    scores, locs = wl.make_head_outputs(3, anchors.shape[0], w2.num_score_cols, 0.0, gen)
    run_pipeline_case("mixup_ignoreband_tiny", w2, gt, anchors, scores, locs, out_dir)

    print("nms cases")
    np.savez_compressed(os.path.join(out_dir, "nms.npz"), **nms_cases(gen))

    print("BASELINE config cases")
    for name, batch, seed in CONFIG_CASES:
        run_config_case(name, batch, seed, out_dir)

    meta = {"torch": torch.__version__, "torchvision": torchvision.__version__,
            "numpy": np.__version__, "reference_root": REF, "seed": 23}
    with open(os.path.join(out_dir, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("done", meta)


if __name__ == "__main__":
    main()
