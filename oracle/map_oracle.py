"""CPU restatement of the reference's detection metric (TEST INFRASTRUCTURE, not a product path).

``detection/metrics/mean_average_precision.py:10-116`` of georgymironov/single_shot_detection:
greedy matching of score-sorted detections to the ground truth of their (image, class) and
per-class average precision (VOC 11-point or area under the precision envelope).  Only ``tests/``
may import this module.  It is pinned by ``tests/golden/map.npz`` -- values produced by the
reference itself (``tests/golden/make_golden_map.py``).

Semantics kept from the reference:
  * predictions [N, 7] = (image id, x1, y1, x2, y2, class, score), processed in descending score
    order (:41-42; the reference's argsort is not stable -- ties are implementation defined, this
    restatement keeps the input order among equal scores);
  * a detection whose class does not occur in its image's ground truth is a false positive (:56-58);
  * otherwise IoU (bf/utils/box_utils.py:83-101, clamped areas) against ALL boxes of that class in
    the image, argmax = first maximum (:60-61); ``value > iou_threshold`` in fp32 (:62);
  * above the threshold: a DIFFICULT box (7th column != 0, only when the rows have one) counts as
    neither (:63), an unmatched box is a true positive and becomes matched, a matched one a false
    positive (:64-68); at or below the threshold: false positive (:69-70);
  * per class over the keys of total_positive (:26-33, difficult boxes do not count): cumulative
    tp / fp, precision = tp / (tp + fp) with a trailing 0, made non-increasing from the right with
    torch.max (NaN propagates), recall = tp / total_positive (:84-99);
  * VOC: recall gets a trailing 1, eleven thresholds torch.arange(0, 1.1, .1), precision at the
    first index whose recall reaches the threshold, mean (:101-105); otherwise recall is framed by
    0 and 1 and AP = sum((r[k+1] - r[k]) * precision[k]) (:106-108);
  * a class with ground truth but no detection gets tp = [0], fp = [1] -> AP 0 (:87-95);
  * mAP = mean over the classes that have (non-difficult) ground truth (:115).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

CLASS_COL = 4
DIFFICULT_COL = 6


def _iou_row(box: torch.Tensor, others: torch.Tensor) -> torch.Tensor:
    """IoU of one corner box against [G, 4] corner boxes, fp32, ops rounded separately."""
    lt = torch.max(box[:2], others[:, :2])
    rb = torch.min(box[2:], others[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    area_a = (box[2] - box[0]).clamp(min=0) * (box[3] - box[1]).clamp(min=0)
    area_b = (others[:, 2] - others[:, 0]).clamp(min=0) * (others[:, 3] - others[:, 1]).clamp(min=0)
    return inter / ((area_a + area_b) - inter)


def match_detections(predictions: torch.Tensor, gts: Sequence[torch.Tensor], iou_threshold: float):
    """-> (order [N] int64, flags [N] int8 in sorted order: 1 = TP, 2 = FP, 0 = neither,
    classes [N] int64 in sorted order, total_positive {class: count})."""
    ignore_difficult = gts[0].size(1) > DIFFICULT_COL
    total_positive: Dict[int, int] = defaultdict(int)
    for gt in gts:
        for row in gt:
            c = int(row[CLASS_COL])
            if not ignore_difficult or float(row[DIFFICULT_COL]) == 0:
                total_positive[c] += 1
            else:
                total_positive[c] += 0          # the key exists only through non-difficult boxes in the reference
    total_positive = {c: n for c, n in total_positive.items() if n > 0}
    order = torch.sort(predictions[:, 6], descending=True, stable=True).indices
    thr = torch.tensor(iou_threshold, dtype=torch.float32)
    flags = np.zeros(predictions.shape[0], dtype=np.int8)
    classes = np.zeros(predictions.shape[0], dtype=np.int64)
    matched = defaultdict(set)
    for k, idx in enumerate(order.tolist()):
        pred = predictions[idx]
        img, c = int(pred[0]), int(pred[5])
        classes[k] = c
        gt = gts[img]
        rows = (gt[:, CLASS_COL].long() == c).nonzero().flatten()
        if rows.numel() == 0:
            flags[k] = 2
            continue
        iou = _iou_row(pred[1:5].float(), gt[rows, 0:4].float())
        value, index = iou.max(dim=0)
        if bool(value > thr):
            g = int(rows[int(index)])
            if not ignore_difficult or float(gt[g, DIFFICULT_COL]) == 0:
                if g not in matched[img]:
                    flags[k] = 1
                    matched[img].add(g)
                else:
                    flags[k] = 2
        else:
            flags[k] = 2
    return order, flags, classes, total_positive


def average_precision(tp_flags: np.ndarray, fp_flags: np.ndarray, total: int, voc: bool) -> float:
    if tp_flags.size:
        tp = torch.from_numpy(np.cumsum(tp_flags).astype(np.float32))
        fp = torch.from_numpy(np.cumsum(fp_flags).astype(np.float32))
    else:
        tp, fp = torch.tensor([0.0]), torch.tensor([1.0])
    precision = torch.cat([tp / (tp + fp), torch.tensor([0.0])])
    for i in reversed(range(1, len(precision))):
        precision[i - 1] = torch.max(precision[i - 1], precision[i])
    recall = tp / total
    if voc:
        recall = torch.cat([recall, torch.tensor([1.0])])
        idx = torch.arange(0, 1.1, 0.1).unsqueeze(0).expand((recall.size(0), 11)) \
            .gt(recall.unsqueeze(1).expand((recall.size(0), 11))).sum(dim=0)
        return float(precision[idx].mean())
    recall = torch.cat([torch.tensor([0.0]), recall, torch.tensor([1.0])])
    return float((recall[1:] - recall[:-1]).dot(precision))


def mean_average_precision(predictions: torch.Tensor, gts: Sequence[torch.Tensor], iou_threshold: float,
                           voc: bool = False) -> Tuple[float, Dict[int, float], np.ndarray]:
    """-> (mAP, {class: AP}, flags in sorted order)."""
    order, flags, classes, total_positive = match_detections(predictions, gts, iou_threshold)
    aps: Dict[int, float] = {}
    for c in sorted(total_positive):
        sel = classes == c
        aps[c] = average_precision((flags[sel] == 1).astype(np.int64), (flags[sel] == 2).astype(np.int64),
                                   total_positive[c], voc)
    return sum(aps.values()) / len(aps), aps, flags
