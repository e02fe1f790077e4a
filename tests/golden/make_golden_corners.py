"""Golden values of the optional API corners, produced by the REFERENCE ITSELF:
``box_utils.nms(soft=True)`` (bf/utils/box_utils.py:145-163), ``Postprocessor`` with a soft-NMS
config (detection/postprocessor.py) and ``box_utils.generalized_iou`` (:104-143).
Build container only: ``python tests/golden/make_golden_corners.py`` -> ``corners.npz``."""
from __future__ import annotations

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

from bf.utils import box_utils as ref_box_utils  # noqa: E402
from detection.box_coder import BoxCoder as RefBoxCoder  # noqa: E402
from detection.postprocessor import Postprocessor as RefPostprocessor  # noqa: E402
from detection import sampler as ref_sampler  # noqa: E402
from detection.losses.multibox_loss import MultiboxLoss as RefMultiboxLoss  # noqa: E402
from detection.target_assigner import TargetAssigner as RefTargetAssigner  # noqa: E402

from single_shot_detection_b200 import workloads as wl  # noqa: E402


def clustered_boxes(gen, n, img=300.0, clusters=6):
    centres = torch.rand((clusters, 2), generator=gen) * img
    which = torch.randint(0, clusters, (n,), generator=gen)
    c = centres[which] + torch.randn((n, 2), generator=gen) * 12
    s = torch.rand((n, 2), generator=gen) * 80 + 20
    return torch.cat([c - s / 2, c + s / 2], dim=1).float()


def main():
    gen = torch.Generator().manual_seed(23)
    blob, n = {}, 0
    # ---- box_utils.nms(soft=True) on one box set ----
    for count, k, thr, sigma in [(1, None, .01, .5), (2, None, .3, .5), (40, None, .05, .5), (90, 100, .2, .5),
                                 (100, 100, .01, .3), (128, None, .4, 1.0), (60, None, .6, .5), (7, None, .0, .5)]:
        boxes = clustered_boxes(gen, count)
        scores = torch.rand((count,), generator=gen)
        if count == 60:
            scores[0] = 0.99               # index 0 survives every decay: the sum-of-indices loop head drops it last
        (bk, sk), picked = ref_box_utils.nms(boxes.clone(), scores.clone(), overlap_threshold=.45, score_threshold=thr,
                                             max_per_class=k, soft=True, sigma=sigma)
        blob[f"soft_boxes_{n}"] = boxes.numpy()
        blob[f"soft_scores_{n}"] = scores.numpy()
        blob[f"soft_cfg_{n}"] = np.array([thr, sigma, -1 if k is None else k], dtype=np.float64)
        blob[f"soft_picked_{n}"] = picked.numpy()
        blob[f"soft_out_{n}"] = torch.cat([bk.reshape(-1, 4), sk.reshape(-1, 1)], dim=1).numpy()
        print("soft", n, count, k, thr, sigma, "->", picked.numel())
        n += 1
    blob["num_soft"] = np.array(n)

    # ---- Postprocessor with soft-NMS on the tiny workloads ----
    m = 0
    for name, batch in [("tiny_voc_b3", 3), ("tiny_sigmoid_b2", 2)]:
        w = wl.WORKLOADS[name]
        anchors, gt, scores, locs = wl.make_inputs(w, seed=31 + m, batch=batch)
        scores = scores * 2.0                                                  # sharper scores: fewer candidates
        coder = RefBoxCoder(w.xy_scale, w.wh_scale)
        post = RefPostprocessor(coder, score_threshold=.2, nms={"max_per_class": 100, "overlap_threshold": .45,
                                                                "soft": True, "sigma": .5},
                                score_converter=w.converter, max_total=40)
        dets = post.postprocess((scores.clone(), locs.clone()), anchors)
        sizes = [d.shape[0] for d in dets]
        blob[f"post_workload_{m}"] = np.array(name)
        blob[f"post_scores_{m}"] = scores.numpy()
        blob[f"post_locs_{m}"] = locs.numpy()
        blob[f"post_det_flat_{m}"] = torch.cat([d.reshape(-1, 6) for d in dets]).numpy()
        blob[f"post_det_off_{m}"] = np.cumsum([0] + sizes).astype(np.int64)
        print("post", m, name, sizes)
        m += 1
    blob["num_post"] = np.array(m)

    # ---- generalized_iou ----
    a = clustered_boxes(gen, 37)
    b = clustered_boxes(gen, 53)
    a[3] = b[5]                                                               # identical boxes: giou = 1
    blob["giou_a"], blob["giou_b"] = a.numpy(), b.numpy()
    blob["giou_cartesian"] = ref_box_utils.generalized_iou(a, b).numpy()
    blob["giou_elementwise"] = ref_box_utils.generalized_iou(a, b[:37], cartesian=False).numpy()
    # ---- MultiboxLoss with GeneralizedIoULoss as the localisation term: values and gradients (autograd) ----
    import functools
    g = 0
    for name, batch in [("tiny_voc_b3", 3), ("tiny_sigmoid_b2", 2)]:
        w = wl.WORKLOADS[name]
        anchors, gt, scores, locs = wl.make_inputs(w, seed=51 + g, batch=batch)
        locs = locs * 3.0                                                        # decoded boxes that really move
        coder = RefBoxCoder(w.xy_scale, w.wh_scale)
        target = RefTargetAssigner(w.matched_threshold, w.unmatched_threshold).encode_ground_truth(gt, anchors)
        if w.converter == "SOFTMAX":
            smp = functools.partial(ref_sampler.hard_negative_mining, negative_per_positive_ratio=w.ratio,
                                    min_negative_per_image=w.min_neg)
            crit = RefMultiboxLoss(smp, coder, {"name": "CrossEntropyLoss"}, {"name": "GeneralizedIoULoss"},
                                   localization_weight=2.0)
        else:
            crit = RefMultiboxLoss(ref_sampler.naive_sampler, coder, {"name": "SigmoidFocalLoss", "gamma": 2.0, "alpha": 0.25},
                                   {"name": "GeneralizedIoULoss"}, localization_weight=2.0)
        s_in = scores.clone().requires_grad_(True)
        l_in = locs.clone().requires_grad_(True)
        loss, class_loss, loc_loss = crit((s_in, l_in), anchors, target.clone())
        loss.backward()
        blob[f"giou_loss_workload_{g}"] = np.array(name)
        blob[f"giou_loss_scores_{g}"] = scores.numpy()
        blob[f"giou_loss_locs_{g}"] = locs.numpy()
        blob[f"giou_loss_target_{g}"] = target.numpy()
        blob[f"giou_loss_values_{g}"] = np.array([float(x.detach()) for x in (loss, class_loss, loc_loss)], dtype=np.float64)
        blob[f"giou_loss_grad_locs_{g}"] = l_in.grad.numpy()
        blob[f"giou_loss_grad_scores_{g}"] = s_in.grad.numpy()
        print("giou loss", g, name, float(loss.detach()), float(loc_loss.detach()), float(l_in.grad.abs().max()))
        g += 1
    blob["num_giou_loss"] = np.array(g)
    np.savez_compressed(os.path.join(HERE, "corners.npz"), **blob)


if __name__ == "__main__":
    main()
