"""The timed region of SURVEY.md §8(d) wired through the reference-shaped classes.

``AnchorPipeline.step`` is what ``detection/init.py:108-135`` (``step_fn``) does around the
model in an eval step, minus the model and the loss arithmetic:

    target = target_assigner.encode_ground_truth(ground_truth, priors)          init.py:114
    mask   = sampler(scores, target[..., 4].long())                             multibox_loss.py:49,58
    to_centroids(target[..., :4], inplace=True); encode_box(..., inplace=True)  multibox_loss.py:81-82
    dets   = postprocessor.postprocess((scores, locs), priors)                  init.py:121-122

``step`` takes what the reference's callers hand over (a Python list of ground-truth tensors, CPU
anchors, prediction tensors) and returns the reference's types.  ``step_device`` is the same
sequence on device-resident inputs with no host synchronisation, and ``capture`` records it into
a CUDA graph (launch-bound configs: ~10 kernels of a few microseconds each).
"""
from __future__ import annotations

import functools
from typing import Dict, NamedTuple, Optional

import torch

from . import _devcache, box_utils, sampler as _sampler
from .box_coder import BoxCoder
from .postprocessor import Postprocessor
from .target_assigner import CLASS_INDEX, LOC_INDEX_END, LOC_INDEX_START, PackedGroundTruth, TargetAssigner, pack_ground_truth


class StepOutput(NamedTuple):
    target: torch.Tensor          # [B, A, 6], box columns encoded in place
    mask: torch.Tensor            # [B, A] bool sampler output
    dets: torch.Tensor            # [B, T, 6] padded detections
    counts: torch.Tensor          # [B] int32 rows used in dets
    det_anchors: torch.Tensor     # [B, T] int32 anchor of every detection row
    status: torch.Tensor          # [4] int32 (see ssd_postprocess)
    stats: Optional[torch.Tensor]  # [B, 4] int32 positives, hard negatives, ignored, detections (with shard)
    shard: Optional[torch.Tensor]  # packed exchange buffer, when requested
    assign_stats: Optional[torch.Tensor] = None   # [B, 4] int32 positives, ignored, NaN boxes, G (ssd_assign_targets)
    mining_stats: Optional[torch.Tensor] = None   # [B, 4] int32 positives, negatives, selected, ties (ssd_hard_negative_mask)


class AnchorPipeline:
    def __init__(self, cfg: Dict):
        """``cfg`` carries the reference's config dict entries (samples/*.py): matched_threshold,
        unmatched_threshold, sampler, ratio, min_neg, xy_scale, wh_scale, eps, score_threshold,
        overlap_threshold, max_per_class, max_total, converter."""
        self.cfg = dict(cfg)
        self.target_assigner = TargetAssigner(cfg["matched_threshold"], cfg["unmatched_threshold"])
        self.box_coder = BoxCoder(cfg["xy_scale"], cfg["wh_scale"], cfg.get("eps", 1e-8))
        fn = getattr(_sampler, cfg["sampler"])               # detection/init.py:90-92
        kwargs = {k: v for k, v in {"negative_per_positive_ratio": cfg.get("ratio"),
                                    "min_negative_per_image": cfg.get("min_neg")}.items()
                  if k in fn.__code__.co_varnames}
        self.sampler = functools.partial(fn, **kwargs)
        self.postprocessor = Postprocessor(
            self.box_coder, cfg["score_threshold"],
            {"max_per_class": cfg["max_per_class"], "overlap_threshold": cfg["overlap_threshold"]},
            score_converter=cfg["converter"], max_total=cfg["max_total"])
        self.fuse_encode = False       # True: one pass for to_centroids+encode (same rounding)
        self._graph = None
        self._side = None

    # -- the reference-facing call -------------------------------------------------------------
    def step(self, ground_truth, anchors, scores, locs):
        """list of GT tensors, anchors [A,4], scores [B,A*C], locs [B,A*4] (host or device) ->
        (target [B,A,6] with encoded boxes, sampled mask [B,A] bool, list of [n_i,6] detections)."""
        device = torch.device("cuda", torch.cuda.current_device())
        scores = scores if scores.is_cuda else scores.to(device, non_blocking=True)
        locs = locs if locs.is_cuda else locs.to(device, non_blocking=True)
        target = self.target_assigner.encode_ground_truth(ground_truth, anchors)
        mask = self._sample_and_encode(target, anchors, scores)
        dets = self.postprocessor.postprocess((scores, locs), anchors)
        return target, mask, dets

    def _sample_and_encode(self, target, anchors, scores):
        batch, num_anchors = target.shape[:2]
        classes = target[..., CLASS_INDEX].long()
        mask = self.sampler(scores.view(batch, num_anchors, -1), classes)
        self._encode_target_boxes(target, anchors)
        return mask

    def _encode_target_boxes(self, target, anchors):
        target_locs = target[..., LOC_INDEX_START:LOC_INDEX_END]
        if self.fuse_encode:
            self.box_coder.encode_corners_(target_locs, anchors)
        else:
            box_utils.to_centroids(target_locs, inplace=True)                  # multibox_loss.py:81
            self.box_coder.encode_box(target_locs, anchors, inplace=True)      # multibox_loss.py:82

    # -- device-resident, sync-free -------------------------------------------------------------
    def step_device(self, packed: PackedGroundTruth, anchors_dev, scores_dev, locs_dev,
                    shard_capacity: Optional[int] = None) -> "StepOutput":
        """``shard_capacity``: also pack (dets, counts, stats) into the one-buffer layout
        ``sharding.all_gather_detections`` exchanges (so the packing is part of the graph).

        The train-side chain (assign -> sampler -> encode) and the post-processor are independent,
        so they are enqueued on two streams (fork / join with events; capturable into one graph
        with two parallel branches)."""
        main = torch.cuda.current_stream()
        side, side2 = self._side_streams()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            target = self.target_assigner.encode_packed(packed, anchors_dev)
            classes = target[..., CLASS_INDEX].long()                      # multibox_loss.py:49
            assigned = torch.cuda.Event()
            assigned.record(side)
            batch, num_anchors = target.shape[:2]
            mask = self.sampler(scores_dev.view(batch, num_anchors, -1), classes)
        with torch.cuda.stream(side2):
            # the box encoding only needs the assignment: a third branch next to sampler and post-processor
            side2.wait_event(assigned)
            self._encode_target_boxes(target, anchors_dev)
        dets, counts, det_anchors, status = self.postprocessor.postprocess_padded((scores_dev, locs_dev), anchors_dev)
        main.wait_stream(side)
        main.wait_stream(side2)
        mining = _sampler.hard_negative_mining.last_stats if self.cfg["sampler"] == "hard_negative_mining" else None
        stats, shard = None, None
        if shard_capacity is not None:
            from . import sharding
            stats = matched_stats(self.target_assigner.last_stats, mining, counts)
            shard = sharding.pack_shard(dets, counts, stats, shard_capacity)
        return StepOutput(target, mask, dets, counts, det_anchors, status, stats, shard,
                          self.target_assigner.last_stats, mining)

    def _side_streams(self):
        if self._side is None:
            self._side = (torch.cuda.Stream(), torch.cuda.Stream())
        return self._side

    def capture(self, packed: PackedGroundTruth, anchors_dev, scores_dev, locs_dev, warmup: int = 2,
                shard_capacity: Optional[int] = None) -> "StepOutput":
        """Record ``step_device`` on these (static) buffers into a CUDA graph; returns the outputs
        the replays will keep overwriting."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                out = self.step_device(packed, anchors_dev, scores_dev, locs_dev, shard_capacity)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.step_device(packed, anchors_dev, scores_dev, locs_dev, shard_capacity)
        self._graph = graph
        return out

    def replay(self):
        self._graph.replay()


def matched_stats(assign_stats: torch.Tensor, mining_stats: Optional[torch.Tensor], counts: torch.Tensor):
    """[B,4] int32 {positives, hard negatives selected, ignored, detections}."""
    out = torch.zeros((assign_stats.shape[0], 4), dtype=torch.int32, device=assign_stats.device)
    out[:, 0] = assign_stats[:, 0]
    if mining_stats is not None:
        out[:, 1] = mining_stats[:, 2]
    out[:, 2] = assign_stats[:, 1]
    out[:, 3] = counts
    return out
