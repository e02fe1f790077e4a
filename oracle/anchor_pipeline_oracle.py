"""CPU oracle for the SSD anchor pipeline.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch CPU tensor ops + numpy) of the algorithm of
georgymironov/single_shot_detection's per-image anchor pipeline.  It exists so that the CUDA
path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package (``single_shot_detection_b200``) never does and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors of its own, so this
oracle is pinned by *outputs of the reference itself*: ``tests/golden/make_golden.py`` imports
the reference modules from ``/root/reference`` (build container only), runs them on seeded
inputs and commits the results under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function here against those fixtures (bit-exact for indices / masks / keep lists /
IoU / xy coding, 1e-6 for the transcendental paths).

Every function cites the reference ``file:line`` it follows (paths relative to the reference
repository root).  The arithmetic keeps the reference's operation ORDER (each fp32 op rounded
separately, no fused multiply-add), because the order decides the bits.

Third-party arithmetic: hard NMS in the reference is ``torchvision.ops.nms``
(``bf/utils/box_utils.py:193``; ``requirements.txt:7`` pins only ``torchvision>=0.3.0``; the
build container has torchvision 0.26.0+cu128).  ``greedy_nms`` below restates its published
CPU algorithm (stable descending sort, unclamped areas, ``inter / (a_i + a_j - inter)`` tested
against the threshold as float-vs-double) and is pinned against the installed torchvision in
``tests/test_oracle_golden.py``.

Tie handling.  Where the reference relies on an unstable sort / unsorted top-k, ties at the
selection boundary are implementation defined (SURVEY.md §7 hard part 3).  The functions here
take ``canonical=True`` to resolve such ties deterministically (lower index wins), which is the
rule the CUDA kernels implement; ``boundary_tie_*`` helpers let tests detect inputs on which the
reference's own answer is not unique.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# target row layout -- detection/target_assigner.py:7-14, bf/datasets/detection_dataset.py:11-17
LOC_LO, LOC_HI = 0, 4
CLS_COL = 4
SCORE_COL = 5
TARGET_COLS = 6
BACKGROUND = 0
IGNORED = -1
# matcher sentinels -- detection/matcher.py:4-5
UNMATCHED_IDX = -2
IGNORED_IDX = -1


# --------------------------------------------------------------------------------------
# box format helpers -- bf/utils/box_utils.py
# --------------------------------------------------------------------------------------
def corners_from_centroids(box: torch.Tensor) -> torch.Tensor:
    """(cx,cy,w,h) -> (x1,y1,x2,y2).  bf/utils/box_utils.py:16-23 (w/2 is exact in fp32)."""
    half = box[..., 2:] / 2
    centre = box[..., :2]
    return torch.cat((centre - half, centre + half), dim=-1)


def centroids_from_corners(box: torch.Tensor, inplace: bool = False) -> Optional[torch.Tensor]:
    """(x1,y1,x2,y2) -> (cx,cy,w,h).  bf/utils/box_utils.py:25-36.

    The two branches round differently: in place it is ``min + (max-min)/2``,
    out of place ``(max+min)/2``.
    """
    if inplace:
        lo, hi = box[..., :2], box[..., 2:]
        hi -= lo
        lo += hi / 2
        return None
    lo, hi = box[..., :2], box[..., 2:]
    return torch.cat(((hi + lo) / 2, hi - lo), dim=-1)


def box_area(corner_box: torch.Tensor) -> torch.Tensor:
    """Clamped area of corner boxes.  bf/utils/box_utils.py:38-46."""
    w = (corner_box[..., 2] - corner_box[..., 0]).clamp(min=0)
    h = (corner_box[..., 3] - corner_box[..., 1]).clamp(min=0)
    return w * h


def pairwise_overlap_box(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Intersection rectangle of every (a_i, b_j) pair, [Na, Nb, 4].  box_utils.py:49-80."""
    top_left = torch.maximum(a[:, None, :2], b[None, :, :2])
    bottom_right = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    return torch.cat((top_left, bottom_right), dim=-1)


def pairwise_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """IoU matrix [Na, Nb] of corner boxes.  bf/utils/box_utils.py:83-101.

    inter / ((area_a + area_b) - inter); every op is a separately rounded fp32 op and
    0/0 stays NaN, exactly as in the reference.
    """
    inter = box_area(pairwise_overlap_box(a, b))
    union = (box_area(a)[:, None] + box_area(b)[None, :]) - inter
    return inter / union


# --------------------------------------------------------------------------------------
# matcher -- detection/matcher.py:33-56
# --------------------------------------------------------------------------------------
def match_anchors(iou: torch.Tensor, matched_threshold: float,
                  unmatched_threshold: Optional[float] = None,
                  force_match: bool = True) -> torch.Tensor:
    """Per-anchor GT index in {-2 (unmatched), -1 (ignored), 0..G-1}; int64 [A].

    Tie rules (all from ATen CPU semantics the reference runs on):
    column max -> lowest GT index; row argmax -> lowest anchor index; several GTs forcing the
    same anchor -> highest GT index wins; comparisons are done in fp32.
    """
    if unmatched_threshold is None:
        unmatched_threshold = matched_threshold
    assert matched_threshold >= unmatched_threshold          # matcher.py:43
    best_iou, gt_of_anchor = iou.max(dim=0)                   # matcher.py:45
    low = best_iou < unmatched_threshold
    mid = (best_iou < matched_threshold) & ~low
    gt_of_anchor = gt_of_anchor.clone()
    gt_of_anchor[low] = UNMATCHED_IDX                         # matcher.py:49
    gt_of_anchor[mid] = IGNORED_IDX                           # matcher.py:50
    if force_match:
        best_anchor = iou.argmax(dim=1)                       # matcher.py:53
        # sequential assignment, last writer wins                   matcher.py:54
        for g, a in enumerate(best_anchor.tolist()):
            gt_of_anchor[a] = g
    return gt_of_anchor


def greedy_bipartite_match(iou: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """detection/matcher.py:7-31 (dead code in the reference; kept for API completeness)."""
    w = iou.clone()
    num_gt, num_anchor = w.shape
    assert bool((w.max(dim=1)[0] > 0).all())
    anchor_of_gt = torch.empty(num_gt, dtype=torch.long)
    for _ in range(num_gt):
        flat = int(w.argmax())
        g, a = divmod(flat, num_anchor)
        anchor_of_gt[g] = a
        w[:, a] = 0
        w[g] = 0
    return torch.arange(num_gt), anchor_of_gt


# --------------------------------------------------------------------------------------
# target assignment -- detection/target_assigner.py:22-63
# --------------------------------------------------------------------------------------
def assign_targets(gt_per_image: Sequence[torch.Tensor], anchors_cxcywh: torch.Tensor,
                   matched_threshold: float, unmatched_threshold: float,
                   return_match: bool = False):
    """target[B, A, 6] = (x1,y1,x2,y2,class,score) per anchor; fp32.

    Background rows are (0,0,0,0, 0, 1), ignored rows (0,0,0,0, -1, -1), matched rows copy
    the GT corner box, class and score.  Images without GT stay background
    (target_assigner.py:43-44).
    """
    num_images = len(gt_per_image)
    num_anchors = anchors_cxcywh.shape[0]
    anchor_corners = corners_from_centroids(anchors_cxcywh)              # :36
    target = torch.zeros((num_images, num_anchors, TARGET_COLS), dtype=torch.float32)
    target[..., CLS_COL] = float(BACKGROUND)                             # :39
    target[..., SCORE_COL] = 1.0                                         # :40
    matches: List[torch.Tensor] = []
    for i, gt in enumerate(gt_per_image):
        if gt.shape[0] == 0:
            matches.append(torch.full((num_anchors,), UNMATCHED_IDX, dtype=torch.long))
            continue
        iou = pairwise_iou(gt[:, LOC_LO:LOC_HI], anchor_corners)         # :47
        idx = match_anchors(iou, matched_threshold, unmatched_threshold) # :49
        hit = idx >= 0
        src = gt[idx[hit]]
        target[i, hit, LOC_LO:LOC_HI] = src[:, LOC_LO:LOC_HI]            # :52
        target[i, hit, CLS_COL] = src[:, CLS_COL]                        # :53
        target[i, hit, SCORE_COL] = src[:, SCORE_COL]                    # :54
        ign = idx == IGNORED_IDX
        target[i, ign, CLS_COL] = float(IGNORED)                         # :57
        target[i, ign, SCORE_COL] = float(IGNORED)                       # :58
        matches.append(idx)
    if return_match:
        return target, matches
    return target


def positive_rows_have_nan(target: torch.Tensor) -> bool:
    """The reference's runtime assert, detection/target_assigner.py:60-61."""
    cls = target[..., CLS_COL]
    pos = (cls != BACKGROUND) & (cls != IGNORED)
    return bool(torch.isnan(target[..., LOC_LO:LOC_HI][pos]).any())


# --------------------------------------------------------------------------------------
# box coder -- detection/box_coder.py
# --------------------------------------------------------------------------------------
def encode_boxes(boxes_cxcywh: torch.Tensor, priors: torch.Tensor, xy_scale: float,
                 wh_scale: float, eps: float = 1e-8, inplace: bool = False) -> torch.Tensor:
    """Centre-size coding against priors.  detection/box_coder.py:13-34.

    in place:     xy = ((xy - p_xy) / p_wh) * xy_scale ; wh = log(wh / p_wh + eps) * wh_scale
    out of place: xy identical                        ; wh = log((wh + eps) / p_wh) * wh_scale
    """
    p = priors.unsqueeze(0)
    if inplace:
        xy, wh = boxes_cxcywh[..., :2], boxes_cxcywh[..., 2:]
        xy -= p[..., :2]
        xy /= p[..., 2:]
        xy *= xy_scale
        wh /= p[..., 2:]
        wh += eps
        wh.log_()
        wh *= wh_scale
        return boxes_cxcywh
    xy = (boxes_cxcywh[..., :2] - p[..., :2]) / p[..., 2:] * xy_scale
    wh = torch.log((boxes_cxcywh[..., 2:] + eps) / p[..., 2:]) * wh_scale
    return torch.cat((xy, wh), dim=-1)


def decode_boxes(locs: torch.Tensor, priors: torch.Tensor, xy_scale: float, wh_scale: float,
                 inplace: bool = False) -> torch.Tensor:
    """Inverse coding.  detection/box_coder.py:37-57.

    out of place (what every caller uses): xy = p_xy + (p_wh * l_xy) / xy_scale ;
                                           wh = p_wh * exp(l_wh / wh_scale)
    in place:                              xy = (l_xy / xy_scale) * p_wh + p_xy ;
                                           wh = exp(l_wh / wh_scale) * p_wh
    """
    p = priors.unsqueeze(0)
    if inplace:
        xy, wh = locs[..., :2], locs[..., 2:]
        xy /= xy_scale
        xy *= p[..., 2:]
        xy += p[..., :2]
        wh /= wh_scale
        wh.exp_()
        wh *= p[..., 2:]
        return locs
    xy = p[..., :2] + p[..., 2:] * locs[..., :2] / xy_scale
    wh = p[..., 2:] * torch.exp(locs[..., 2:] / wh_scale)
    return torch.cat((xy, wh), dim=-1)


# --------------------------------------------------------------------------------------
# samplers -- detection/sampler.py
# --------------------------------------------------------------------------------------
def positives_mask(target_classes: torch.Tensor) -> torch.Tensor:
    """detection/sampler.py:9-10."""
    return (target_classes != BACKGROUND) & (target_classes != IGNORED)


def background_loss(logits: torch.Tensor) -> torch.Tensor:
    """-log_softmax(logits)[..., 0], the mining criterion.  detection/sampler.py:13."""
    return -F.log_softmax(logits, dim=-1)[:, :, BACKGROUND]


def negatives_to_keep(target_classes: torch.Tensor, ratio, min_per_image) -> torch.Tensor:
    """min(max(n_pos * ratio, min_per_image), n_neg) per image, [B,1].  sampler.py:15-20."""
    n_neg = (target_classes == BACKGROUND).sum(dim=1, keepdim=True)
    n_pos = positives_mask(target_classes).sum(dim=1, keepdim=True)
    return torch.min(torch.clamp(n_pos * ratio, min=min_per_image), n_neg)


def mine_hard_negatives(logits: torch.Tensor, target_classes: torch.Tensor, ratio,
                        min_per_image, canonical: bool = True,
                        loss: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3:1 online hard-negative mining mask, bool [B, A].  detection/sampler.py:12-25.

    ``canonical=False`` ranks with the reference's double (unstable) argsort; ``canonical=True``
    ranks by (loss descending, anchor index ascending), which selects the same set whenever no
    loss value ties across the cut.  ``loss`` lets a caller inject the mining criterion
    (stage-boundary parity: identical fp32 inputs to the selection).
    """
    if loss is None:
        loss = background_loss(logits)
    else:
        loss = loss.clone()
    neg = target_classes == BACKGROUND
    pos = positives_mask(target_classes)
    keep_n = negatives_to_keep(target_classes, ratio, min_per_image)
    loss[~neg] = -math.inf                                               # sampler.py:21
    if canonical:
        order = torch.sort(loss, dim=1, descending=True, stable=True)[1]
        rank = torch.empty_like(order)
        rank.scatter_(1, order, torch.arange(loss.shape[1]).expand_as(order))
    else:
        rank = loss.argsort(dim=1, descending=True).argsort(dim=1)       # sampler.py:22
    return pos | (rank < keep_n)                                         # sampler.py:23-25


def mining_boundary_tie(loss: torch.Tensor, target_classes: torch.Tensor, ratio,
                        min_per_image) -> torch.Tensor:
    """bool [B]: True where the k-th and (k+1)-th largest negative losses are equal, i.e. the
    reference's own selection is not unique for that image."""
    neg = target_classes == BACKGROUND
    keep_n = negatives_to_keep(target_classes, ratio, min_per_image).view(-1)
    masked = loss.clone()
    masked[~neg] = -math.inf
    ordered = torch.sort(masked, dim=1, descending=True)[0]
    out = torch.zeros(loss.shape[0], dtype=torch.bool)
    for i, k in enumerate(keep_n.tolist()):
        k = int(math.ceil(k))
        if 0 < k < int(neg[i].sum()):
            out[i] = bool(ordered[i, k - 1] == ordered[i, k])
    return out


# --------------------------------------------------------------------------------------
# NMS -- bf/utils/box_utils.py:165-194 and torchvision.ops.nms (CPU kernel)
# --------------------------------------------------------------------------------------
def greedy_nms(corner_boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """Restatement of torchvision's CPU nms: returns kept indices in descending-score order.

    * candidates visited in stable descending score order;
    * area = (x2-x1)*(y2-y1), NOT clamped; intersection sides clamped at 0;
    * ovr = inter / (area_i + area_j - inter) in fp32, suppressed when ovr > threshold with the
      comparison done in double (the threshold is a C++ double).
    """
    b = np.ascontiguousarray(corner_boxes, dtype=np.float32)
    s = np.ascontiguousarray(scores, dtype=np.float32)
    n = b.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-s, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    thr = np.float64(iou_threshold)
    kept = []
    zero = np.float32(0)
    with np.errstate(divide="ignore", invalid="ignore"):
        for pos in range(n):
            i = order[pos]
            if dead[i]:
                continue
            kept.append(i)
            rest = order[pos + 1:]
            if rest.size == 0:
                continue
            w = np.maximum(zero, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
            h = np.maximum(zero, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
            inter = w * h
            ovr = inter / ((area[i] + area[rest]) - inter)
            dead[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(kept, dtype=np.int64)


def gaussian_soft_nms(corner_boxes: torch.Tensor, scores: torch.Tensor, score_threshold: float,
                      sigma: float = 0.5) -> torch.Tensor:
    """Soft-NMS exactly as bf/utils/box_utils.py:145-163 runs it; returns the picked indices in pick
    order (the reference returns ``(boxes[picked], scores[picked])`` with the INPUT scores).

    Kept quirks: the loop head tests ``mask.nonzero().sum()`` -- the SUM OF THE INDICES of the mask,
    so a mask whose only element is index 0 ends the loop (:151); the mask tested by the next loop head
    is computed before the decay of this iteration (:156 vs :160); ``argmax`` runs over every
    remaining score, above the threshold or not (:152).  Boxes areas are clamped (box_utils.area).
    """
    work = scores.clone()
    mask = scores > score_threshold                                           # :147
    area = box_area(corner_boxes)                                             # :148
    picked: List[int] = []
    while int(mask.nonzero().sum()) and len(picked) < scores.shape[0]:
        idx = int(work.argmax())                                              # :152
        work[idx] = 0                                                         # :153
        picked.append(idx)
        mask = work > score_threshold                                         # :156
        inter_box = pairwise_overlap_box(corner_boxes[idx].unsqueeze(0), corner_boxes[mask]).squeeze(0)
        inter = box_area(inter_box)                                           # :158
        iou = inter / (area[idx] + area[mask] - inter)                        # :159
        work[mask] = work[mask] * iou.pow(2).div_(sigma).neg_().exp_()        # :160
    return torch.tensor(picked, dtype=torch.long)


def generalized_iou(a: torch.Tensor, b: torch.Tensor, cartesian: bool = True) -> torch.Tensor:
    """bf/utils/box_utils.py:104-143: iou - (enclosing - union) / enclosing for corner boxes."""
    if cartesian:
        inter = box_area(pairwise_overlap_box(a, b))
        area_a = box_area(a).unsqueeze(1).expand_as(inter)
        area_b = box_area(b).unsqueeze(0).expand_as(inter)
        lo = torch.min(a[:, None, :2].expand(a.shape[0], b.shape[0], 2), b[None, :, :2].expand(a.shape[0], b.shape[0], 2))
        hi = torch.max(a[:, None, 2:].expand(a.shape[0], b.shape[0], 2), b[None, :, 2:].expand(a.shape[0], b.shape[0], 2))
    else:
        inter = box_area(torch.cat([torch.max(a[..., :2], b[..., :2]), torch.min(a[..., 2:], b[..., 2:])], dim=-1))
        area_a, area_b = box_area(a), box_area(b)
        lo, hi = torch.min(a[..., :2], b[..., :2]), torch.max(a[..., 2:], b[..., 2:])
    union = area_a + area_b - inter
    enclosing = box_area(torch.cat([lo, hi], dim=-1))
    return inter / union - (enclosing - union) / enclosing


def select_top_scores(scores: torch.Tensor, k: int, canonical: bool = True) -> torch.Tensor:
    """Indices of the k largest scores.  Canonical: (score desc, index asc), returned in that
    order; otherwise ``torch.topk(sorted=False)`` as bf/utils/box_utils.py:186-187."""
    if canonical:
        order = torch.sort(scores, descending=True, stable=True)[1]
        return order[:k]
    return scores.topk(k, sorted=False, largest=True)[1]


def class_nms(corner_boxes: torch.Tensor, scores: torch.Tensor, overlap_threshold: float,
              max_per_class: Optional[int] = None, canonical: bool = True,
              use_torchvision: bool = False, soft: Optional[Tuple[float, float]] = None):
    """Top-k then hard NMS for one class (bf/utils/box_utils.py:165-194), or soft-NMS when
    ``soft = (score_threshold, sigma)``.

    Returns ((boxes[keep], scores[keep]), keep, subset) where ``keep`` indexes the top-k
    ``subset`` (as in the reference) and ``subset`` maps back to the input rows (None when no
    top-k was taken).
    """
    subset = None
    if max_per_class is not None and max_per_class < scores.shape[0]:
        subset = select_top_scores(scores, max_per_class, canonical)
        scores = scores[subset]
        corner_boxes = corner_boxes[subset]
    if soft is not None:
        keep = gaussian_soft_nms(corner_boxes, scores, soft[0], soft[1])
    elif use_torchvision:
        import torchvision
        keep = torchvision.ops.nms(corner_boxes, scores, overlap_threshold)
    else:
        keep = torch.from_numpy(greedy_nms(corner_boxes.numpy(), scores.numpy(), overlap_threshold))
    return (corner_boxes[keep], scores[keep]), keep, subset


# --------------------------------------------------------------------------------------
# post-processor -- detection/postprocessor.py:24-78
# --------------------------------------------------------------------------------------
def convert_scores(logits: torch.Tensor, converter: str) -> torch.Tensor:
    """[B, A, C] logits -> foreground probabilities.  postprocessor.py:16-22, 42-48.

    SOFTMAX drops column 0 (background); SIGMOID keeps every column.
    """
    logits = logits.float()
    if converter == "SOFTMAX":
        return F.softmax(logits, dim=-1)[..., 1:]
    if converter == "SIGMOID":
        return torch.sigmoid(logits)
    raise ValueError(f"Wrong value for score_converter: {converter}")


def detections_from_scores(fg_probs: torch.Tensor, corner_boxes: torch.Tensor,
                           score_threshold: float, overlap_threshold: float,
                           max_per_class: Optional[int], max_total: Optional[int],
                           canonical: bool = True, use_torchvision: bool = False,
                           return_keep: bool = False, soft_sigma: Optional[float] = None):
    """Selection half of the post-processor (postprocessor.py:57-76): per class threshold ->
    top-k -> NMS -> concatenate -> optional final top-k.  Inputs are probabilities [B, A, Cf] and
    decoded corner boxes [B, A, 4], so a test can feed it the exact fp32 values another
    implementation produced ("identical stage inputs").

    Returns a list of [n_i, 6] (x1,y1,x2,y2,class,score).  With ``return_keep`` also returns,
    per image and class, the kept ANCHOR indices in keep order (descending score).
    """
    out = []
    kept_anchors = []
    num_fg = fg_probs.shape[-1]
    for probs, boxes in zip(fg_probs, corner_boxes):
        rows = []
        per_class = []
        for c in range(num_fg):
            col = probs[:, c]
            above = col > score_threshold                               # :62 (fp32 compare)
            (b_keep, s_keep), keep, subset = class_nms(boxes[above], col[above], overlap_threshold,
                                                       max_per_class, canonical, use_torchvision,
                                                       None if soft_sigma is None else (score_threshold, soft_sigma))
            if return_keep:             # bookkeeping the reference does not do: only when asked for
                cand = torch.nonzero(above).view(-1)
                src = cand if subset is None else cand[subset]
                per_class.append(src[keep])
            cls = torch.full((s_keep.shape[0], 1), float(c + 1))        # :66
            rows.append(torch.cat((b_keep, cls, s_keep.unsqueeze(1)), dim=-1))
        picked = torch.cat(rows, dim=0)
        if max_total is not None and max_total < picked.shape[0]:       # :72-74
            if canonical:
                top = torch.sort(picked[:, 5], descending=True, stable=True)[1][:max_total]
            else:
                top = torch.topk(picked[:, 5], max_total, sorted=True, largest=True)[1]
            picked = picked[top]
        out.append(picked)
        kept_anchors.append(per_class)
    if return_keep:
        return out, kept_anchors
    return out


def postprocess(scores: torch.Tensor, locs: torch.Tensor, priors: torch.Tensor, *,
                xy_scale: float, wh_scale: float, score_threshold: float,
                overlap_threshold: float, max_per_class: Optional[int],
                max_total: Optional[int], converter: str = "SOFTMAX",
                canonical: bool = True, use_torchvision: bool = False, soft_sigma: Optional[float] = None):
    """Full post-processor.  detection/postprocessor.py:24-78."""
    batch = scores.shape[0]
    num_priors = priors.shape[0]
    probs = convert_scores(scores.view(batch, num_priors, -1), converter)
    decoded = decode_boxes(locs.float().view(batch, num_priors, 4), priors, xy_scale, wh_scale)
    corners = corners_from_centroids(decoded)
    return detections_from_scores(probs, corners, score_threshold, overlap_threshold,
                                  max_per_class, max_total, canonical, use_torchvision, soft_sigma=soft_sigma)


def class_topk_boundary_tie(fg_probs: torch.Tensor, score_threshold: float,
                            max_per_class: int) -> torch.Tensor:
    """bool [B, Cf]: True where the k-th and (k+1)-th best candidate scores of a class tie
    (the reference's ``topk(sorted=False)`` answer is then not unique)."""
    b, _, cf = fg_probs.shape
    out = torch.zeros((b, cf), dtype=torch.bool)
    masked = torch.where(fg_probs > score_threshold, fg_probs, torch.full_like(fg_probs, -1.0))
    if masked.shape[1] <= max_per_class:
        return out
    top = masked.topk(max_per_class + 1, dim=1, sorted=True)[0]       # [B, k+1, Cf]
    kth, nxt = top[:, max_per_class - 1, :], top[:, max_per_class, :]
    return (kth == nxt) & (nxt > 0)


# --------------------------------------------------------------------------------------
# the caller's loss (detection/losses/multibox_loss.py:35-94) for the <=1e-5 loss check
# --------------------------------------------------------------------------------------
def multibox_loss_ce_smoothl1(scores: torch.Tensor, locs: torch.Tensor, priors: torch.Tensor,
                              target: torch.Tensor, sampled_mask: torch.Tensor,
                              encoded_target_locs: torch.Tensor):
    """CrossEntropy + SmoothL1 loss triple as MultiboxLoss.forward computes it when given a
    sampler mask and already encoded target boxes (multibox_loss.py:56-92)."""
    b, a = target.shape[:2]
    cls = target[..., CLS_COL].long()
    pos = positives_mask(cls)
    logits = scores.view(b, a, -1)[sampled_mask]
    class_loss = F.cross_entropy(logits, cls[sampled_mask].view(-1), ignore_index=IGNORED,
                                 reduction="sum")
    loc_loss = F.smooth_l1_loss(locs.view(b, a, 4)[pos].view(-1, 4),
                                encoded_target_locs[pos].view(-1, 4), reduction="sum")
    div = pos.sum().clamp(min=1).float()
    class_loss = class_loss / div
    loc_loss = loc_loss / div
    return class_loss + loc_loss, class_loss, loc_loss


def sigmoid_focal_loss_rows(prediction: torch.Tensor, target: torch.Tensor, gamma: float, alpha: float):
    """Per-row SigmoidFocalLoss terms, bf/modules/losses.py:42-52 (before the reduction)."""
    alpha_weight = target * alpha + (1.0 - target) * (1.0 - alpha)
    pb = torch.sigmoid(prediction)
    pb = pb * target + (1.0 - pb) * (1.0 - target)
    cross_entropy = F.binary_cross_entropy_with_logits(prediction, target, reduction="none")
    return (alpha_weight * (1.0 - pb).pow(gamma) * cross_entropy).sum(dim=-1)


def multibox_loss_focal_smoothl1(scores: torch.Tensor, locs: torch.Tensor, target: torch.Tensor,
                                 sampled_mask: torch.Tensor, encoded_target_locs: torch.Tensor,
                                 gamma: float = 2.0, alpha: float = 0.25):
    """SigmoidFocal + SmoothL1 triple: the MULTICLASS branch of MultiboxLoss.forward
    (multibox_loss.py:60-67): one-hot target at column class-1 scaled by the GT score.

    The classification term is the MEAN over the sampled anchors, not the sum: MultiboxLoss passes
    reduction='sum' (multibox_loss.py:24), but get_ctor wraps the constructor in filter_kwargs
    (bf/utils/misc_utils.py:21-29), which keeps only keywords NAMED in the signature --
    SigmoidFocalLoss.__init__(self, gamma, alpha, **kwargs) names neither `reduction` nor
    `ignore_index`, so _Loss falls back to its default reduction='mean' (losses.py:11).  The golden
    fixtures (the reference's own output) pin this; an empty mask gives NaN as torch's mean does."""
    b, a = target.shape[:2]
    cls = target[..., CLS_COL].long()
    pos = positives_mask(cls)
    logits = scores.view(b, a, -1)[sampled_mask]
    cls_s = cls[sampled_mask]
    score_s = target[..., SCORE_COL][sampled_mask]
    class_target = torch.zeros_like(logits)
    m = positives_mask(cls_s)
    class_target[m, cls_s[m] - 1] = score_s[m]
    class_loss = sigmoid_focal_loss_rows(logits, class_target, gamma, alpha).mean()
    loc_loss = F.smooth_l1_loss(locs.view(b, a, 4)[pos].view(-1, 4),
                                encoded_target_locs[pos].view(-1, 4), reduction="sum")
    div = pos.sum().clamp(min=1).float()
    class_loss = class_loss / div
    loc_loss = loc_loss / div
    return class_loss + loc_loss, class_loss, loc_loss


def giou_localization_loss(locs: torch.Tensor, priors: torch.Tensor, target: torch.Tensor, xy_scale: float,
                           wh_scale: float, loc_weight: float = 1.0):
    """The IOU_LOSS branch of MultiboxLoss.forward (multibox_loss.py:77-79, 84-90) with GeneralizedIoULoss
    (bf/modules/losses.py:109-114, reduction='sum' -- `reduction` IS in its constructor's signature, so
    filter_kwargs keeps it): the predictions are decoded and turned into corner boxes, the target stays a
    corner box; sum over the positives of 1 - giou, times the weight, over max(#positives, 1).
    ``locs`` may require grad (torch autograd gives the reference's gradient)."""
    b, a = target.shape[:2]
    pos = positives_mask(target[..., CLS_COL].long())
    corners = corners_from_centroids(decode_boxes(locs.view(b, a, 4), priors, xy_scale, wh_scale))
    loss = (1.0 - generalized_iou(corners[pos].view(-1, 4), target[..., :4][pos].view(-1, 4), cartesian=False)).sum()
    return loss * loc_weight / pos.sum().clamp(min=1).float()


# --------------------------------------------------------------------------------------
# whole timed region (SURVEY.md §8 d) as the reference executes it on the CPU
# --------------------------------------------------------------------------------------
def run_step(gt_per_image, anchors, scores, locs, cfg, *, canonical: bool = False,
             use_torchvision: bool = True, stage_seconds: Optional[dict] = None):
    """assign -> sampler -> to_centroids+encode (in place) -> postprocess, CPU.

    ``cfg`` keys: matched_threshold, unmatched_threshold, sampler ('hard_negative_mining' |
    'naive_sampler'), ratio, min_neg, xy_scale, wh_scale, eps, score_threshold,
    overlap_threshold, max_per_class, max_total, converter.
    With the defaults (reference tie behaviour, torchvision NMS) this executes the same torch
    CPU ops in the same order as the reference -- including the NaN assertion that closes
    target_assigner.py:60-61 -- and is what ``bench.py`` times as the CPU arm.  ``stage_seconds``
    (a dict) accumulates the wall time of the four stages of SURVEY.md 8(d).
    """
    import time
    clock = time.perf_counter
    b = len(gt_per_image)
    a = anchors.shape[0]
    t0 = clock()
    target = assign_targets(gt_per_image, anchors, cfg["matched_threshold"],
                            cfg["unmatched_threshold"])
    assert not positive_rows_have_nan(target)                            # target_assigner.py:60-61
    t1 = clock()
    cls = target[..., CLS_COL].long()                                    # multibox_loss.py:49
    logits = scores.view(b, a, -1)
    if cfg["sampler"] == "hard_negative_mining":
        mask = mine_hard_negatives(logits, cls, cfg["ratio"], cfg["min_neg"], canonical=canonical)
    else:
        mask = positives_mask(cls)
    t2 = clock()
    tl = target[..., LOC_LO:LOC_HI]
    centroids_from_corners(tl, inplace=True)
    encode_boxes(tl, anchors, cfg["xy_scale"], cfg["wh_scale"], cfg.get("eps", 1e-8), inplace=True)
    t3 = clock()
    dets = postprocess(scores, locs, anchors, xy_scale=cfg["xy_scale"], wh_scale=cfg["wh_scale"],
                       score_threshold=cfg["score_threshold"],
                       overlap_threshold=cfg["overlap_threshold"],
                       max_per_class=cfg["max_per_class"], max_total=cfg["max_total"],
                       converter=cfg["converter"], canonical=canonical,
                       use_torchvision=use_torchvision)
    t4 = clock()
    if stage_seconds is not None:
        for key, dt in (("encode_ground_truth", t1 - t0), ("sampler", t2 - t1), ("to_centroids+encode_box", t3 - t2),
                        ("postprocess", t4 - t3)):
            stage_seconds[key] = stage_seconds.get(key, 0.0) + dt
    return target, mask, dets
