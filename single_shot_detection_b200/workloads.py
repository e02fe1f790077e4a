"""Named workloads (BASELINE.json configs) and their synthetic inputs.

Everything here is input preparation on the host: anchor tables of the right shape, ragged
ground-truth lists, random-init head outputs.  The recipe is SURVEY.md §8(d) / BASELINE.md §2:
seed 23 (``seed = 23`` in every reference ``samples/*.py:1``), GT count ~ UniformInt[1, Gmax],
centres U(0,W)^2, sides exp(U(log .05W, log .6W)), clipped to [0, W-1], zero-size rows dropped
(``bf/datasets/detection_dataset.py:31-32``), class uniform over the foreground ids, score 1.
Head outputs: logits N(0,1) (N(-4.6,1) for the sigmoid head, ``retina_rn50_500_coco.py:27``),
locs N(0, 0.1^2).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import anchors as _anchors

_SSD6 = [[1.0, 2.0]] + [[1.0, 2.0, 3.0]] * 3 + [[1.0, 2.0]] * 2
_SSD7 = [[1.0, 2.0]] + [[1.0, 2.0, 3.0]] * 4 + [[1.0, 2.0]] * 2


@dataclasses.dataclass(frozen=True)
class Workload:
    name: str
    img: int                       # square input side in pixels
    fmaps: Tuple[int, ...]         # square feature-map sides
    anchor_kind: str               # 'ssd' | 'retina'
    anchor_args: Tuple             # see build_anchors
    num_score_cols: int            # C (softmax: incl. background col 0; sigmoid: fg only)
    batch: int
    converter: str                 # 'SOFTMAX' | 'SIGMOID'
    sampler: str                   # 'hard_negative_mining' | 'naive_sampler'
    matched_threshold: float
    unmatched_threshold: float
    overlap_threshold: float
    max_gt: int
    logit_mean: float = 0.0
    score_threshold: float = 0.01
    max_per_class: int = 100
    max_total: int = 200
    xy_scale: float = 10.0
    wh_scale: float = 5.0
    eps: float = 1e-8
    ratio: int = 3
    min_neg: int = 5

    @property
    def num_fg(self) -> int:
        return self.num_score_cols - 1 if self.converter == "SOFTMAX" else self.num_score_cols

    def cfg(self) -> Dict:
        return dict(matched_threshold=self.matched_threshold,
                    unmatched_threshold=self.unmatched_threshold, sampler=self.sampler,
                    ratio=self.ratio, min_neg=self.min_neg, xy_scale=self.xy_scale,
                    wh_scale=self.wh_scale, eps=self.eps, score_threshold=self.score_threshold,
                    overlap_threshold=self.overlap_threshold, max_per_class=self.max_per_class,
                    max_total=self.max_total, converter=self.converter)


def _ssd(name, img, fmaps, lo, hi, ratios, c, batch, max_gt) -> Workload:
    return Workload(name, img, tuple(fmaps), "ssd", (lo, hi, tuple(tuple(r) for r in ratios)), c,
                    batch, "SOFTMAX", "hard_negative_mining", 0.5, 0.5, 0.45, max_gt)


WORKLOADS: Dict[str, Workload] = {w.name: w for w in [
    # BASELINE.json configs[0]: SSD300-VGG16 VOC, 8732 anchors, 21 classes, batch 8 (CPU-sized)
    _ssd("ssd300_voc_b8", 300, (38, 19, 10, 5, 3, 1), 0.15, 1.05, _SSD6, 21, 8, 16),
    # the metric line: SSD300 b32
    _ssd("ssd300_voc_b32", 300, (38, 19, 10, 5, 3, 1), 0.15, 1.05, _SSD6, 21, 32, 16),
    # the reference's own builder yields 37/18/9/5/3/2 maps -> 8108 anchors
    _ssd("ssd300_voc_8108_b8", 300, (37, 18, 9, 5, 3, 2), 0.15, 1.05, _SSD6, 21, 8, 16),
    # configs[1]: SSD-MobileNetV2 COCO, 2268 anchors, 81 classes, batch 64
    _ssd("ssd_mb2_coco_b64", 300, (19, 10, 5, 3, 2, 1), 0.1, 1.05, _SSD6, 81, 64, 32),
    # configs[2]: SSD512-VGG16 COCO, 24564 anchors, batch 32
    _ssd("ssd512_coco_b32", 512, (64, 32, 16, 8, 4, 2, 1), 0.1, 1.05, _SSD7, 81, 32, 32),
    # configs[3]: RetinaNet-R50-500 COCO, 47961 anchors, sigmoid, naive sampler
    Workload("retina500_coco_b32", 500, (63, 32, 16, 8, 4), "retina", ((1.0, 2.0, 0.5), 3, 4.0, 3),
             80, 32, "SIGMOID", "naive_sampler", 0.5, 0.4, 0.5, 32, logit_mean=-4.6),
    # configs[4]: M2Det-512-VGG16 COCO, 24528 anchors, batch 256 (sharded over GPUs)
    _ssd("m2det512_coco_b256", 512, (64, 32, 16, 8, 4, 2), 0.07, 1.05, _SSD6, 81, 256, 32),
    # small cases for tests
    _ssd("tiny_voc_b3", 64, (8, 4, 2, 1), 0.2, 0.9, [[1.0, 2.0]] * 4, 6, 3, 5),
    Workload("tiny_sigmoid_b2", 64, (8, 4), "retina", ((1.0, 2.0, 0.5), 3, 4.0, 2), 4, 2,
             "SIGMOID", "naive_sampler", 0.5, 0.4, 0.5, 4, logit_mean=-3.0),
]}

# the configuration BASELINE.json's metric is quoted on
HEADLINE = "ssd300_voc_b32"


def build_anchors(w: Workload) -> torch.Tensor:
    sizes = [(s, s) for s in w.fmaps]
    if w.anchor_kind == "ssd":
        lo, hi, ratios = w.anchor_args
        return _anchors.ssd_anchor_table((w.img, w.img), sizes, lo, hi, [list(r) for r in ratios])
    ratios, min_level, scale, spl = w.anchor_args
    return _anchors.retina_anchor_table((w.img, w.img), sizes, list(ratios), min_level, scale, spl)


def build_anchor_generators(w: Workload):
    """The workload's per-level generators (anchor_generators.py: the table is written on the GPU)."""
    from . import anchor_generators as ag
    if w.anchor_kind == "ssd":
        lo, hi, ratios = w.anchor_args
        return ag.build_ssd_anchor_generators(num_scales=len(w.fmaps), min_scale=lo, max_scale=hi,
                                              aspect_ratios=[list(r) for r in ratios])
    ratios, min_level, scale, spl = w.anchor_args
    return ag.build_retina_anchor_generators(list(ratios), min_level, min_level + len(w.fmaps) - 1, scale, spl)


def make_ground_truth(batch: int, img: int, num_fg: int, max_gt: int, gen: torch.Generator,
                      extra_col: bool = False, mixup: Optional[float] = None) -> List[torch.Tensor]:
    """List of [G_i, 6(+1)] fp32 rows (x1,y1,x2,y2,class,score[,difficult])."""
    out = []
    lo, hi = math.log(0.05 * img), math.log(0.6 * img)
    for _ in range(batch):
        g = int(torch.randint(1, max_gt + 1, (1,), generator=gen))
        centre = torch.rand((g, 2), generator=gen) * img
        side = torch.exp(torch.rand((g, 2), generator=gen) * (hi - lo) + lo)
        box = torch.cat([centre - side / 2, centre + side / 2], dim=1).clamp_(0, img - 1)
        keep = (box[:, 0] != box[:, 2]) & (box[:, 1] != box[:, 3])
        box = box[keep]
        cls = torch.randint(1, num_fg + 1, (box.shape[0], 1), generator=gen).float()
        score = torch.ones((box.shape[0], 1))
        cols = [box, cls, score]
        if extra_col:
            cols.append(torch.zeros((box.shape[0], 1)))
        rows = torch.cat(cols, dim=1).float()
        if mixup is not None and rows.shape[0] > 0:
            # bf/core/batch_container.py:25-45: rows duplicated with scores lam / 1-lam
            a, b = rows.clone(), rows.clone()
            a[:, 5] *= mixup
            b[:, 5] *= 1.0 - mixup
            rows = torch.cat([a, b], dim=0)
        out.append(rows.contiguous())
    return out


def make_head_outputs(batch: int, num_anchors: int, num_cols: int, logit_mean: float,
                      gen: torch.Generator) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scores [B, A*C], locs [B, A*4]) fp32, the layout ``detection/detector.py:50-66`` emits."""
    scores = torch.randn((batch, num_anchors * num_cols), generator=gen) + logit_mean
    locs = torch.randn((batch, num_anchors * 4), generator=gen) * 0.1
    return scores, locs


def make_inputs(w: Workload, seed: int = 23, batch: Optional[int] = None):
    """(anchors [A,4], gt list, scores [B,A*C], locs [B,A*4]) on the CPU."""
    gen = torch.Generator().manual_seed(seed)
    b = w.batch if batch is None else batch
    anchors = build_anchors(w)
    gt = make_ground_truth(b, w.img, w.num_fg, w.max_gt, gen)
    scores, locs = make_head_outputs(b, anchors.shape[0], w.num_score_cols, w.logit_mean, gen)
    return anchors, gt, scores, locs


def algorithmic_bytes_per_image(w: Workload, num_anchors: int, batch: int) -> float:
    """every input read once, every output written once per API call."""
    a, c = num_anchors, w.num_score_cols
    logits_reads = 2 if w.sampler == "hard_negative_mining" else 1
    return logits_reads * 4.0 * a * c + 77.0 * a + 32.0 * a / batch + 4800.0
