"""TargetAssigner -- same interface as the reference's ``detection/target_assigner.py``.

``encode_ground_truth(ground_truth, anchors)`` takes the list of per-image ground-truth tensors
``[G_i, >=6]`` (x1,y1,x2,y2,class,score[,difficult], pixels) and the anchors ``[A,4]``
(cx,cy,w,h) and returns ``target[B,A,6]`` fp32.  Unlike the reference -- which runs the whole
thing on ``anchors.device``, i.e. on the CPU (detection/anchor_generators/_anchor_generator.py:4,18)
-- the result lives on the GPU: the caller's ``target.to(device)`` (detection/init.py:115) becomes
a no-op.  One launch does IoU, both argmaxes, the forced match and the target write for the whole
batch (csrc/assign.cu).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _devcache
from .ops import OPS

# detection/target_assigner.py:7-14
LOC_INDEX_START = 0
LOC_INDEX_END = 4
CLASS_INDEX = 4
SCORE_INDEX = 5
TARGET_SIZE = 6

NEGATIVE_CLASS = 0
IGNORE_CLASS = -1


class PackedGroundTruth:
    """Ground truth of a batch as the kernel wants it: rows ``[sum G_i, cols]`` + int32 offsets
    ``[B+1]`` on the device.  Build once with :func:`pack_ground_truth` when the same batch is
    assigned repeatedly (benchmarks, CUDA-graph replay)."""

    def __init__(self, rows: torch.Tensor, offsets: torch.Tensor, max_gt: int, batch: int):
        self.rows, self.offsets, self.max_gt, self.batch = rows, offsets, max_gt, batch


def pack_ground_truth(ground_truth: Sequence[torch.Tensor], device: torch.device,
                       out: Optional[PackedGroundTruth] = None) -> PackedGroundTruth:
    """CSR-pack the list into ONE pinned staging buffer and ship it with one H2D copy.

    ``out``: static device buffers to fill instead (CUDA-graph replay: ``out.rows`` [capacity, cols],
    ``out.offsets`` [B+1], ``out.max_gt`` the capacity per image the graph was captured with);
    raises ValueError when the batch does not fit."""
    batch = len(ground_truth)
    sizes = [int(g.shape[0]) for g in ground_truth]
    cols = min([int(g.shape[1]) for g in ground_truth if g.dim() == 2 and g.shape[0]] or [TARGET_SIZE])
    if cols < TARGET_SIZE:
        raise ValueError(f"ground-truth rows need at least {TARGET_SIZE} columns, got {cols}")
    total = sum(sizes)
    head = (batch + 1 + 3) // 4 * 4                      # int32 offsets, padded to 16 bytes
    stage = _devcache.pinned_words(head + total * cols)
    off = stage[: batch + 1].view(torch.int32)
    acc = 0
    offs = [0]
    for s in sizes:
        acc += s
        offs.append(acc)
    off.copy_(torch.tensor(offs, dtype=torch.int32))
    rows_host = stage[head: head + total * cols].view(total, cols)
    if total:
        on_device = [g for g in ground_truth if g.shape[0] and g.is_cuda]
        if on_device:                                     # already on the GPU: gather there
            parts = [g[:, :cols].to(device=device, dtype=torch.float32) for g in ground_truth if g.shape[0]]
            rows_dev = torch.cat(parts, dim=0).contiguous()
            offsets_dev = off.to(device, non_blocking=True)
            _devcache.mark_in_flight()
            return PackedGroundTruth(rows_dev, offsets_dev, max(sizes), batch)
        torch.cat([g[:, :cols] for g in ground_truth if g.shape[0]], dim=0, out=rows_host)
    if out is not None:
        if (batch != out.batch or total > out.rows.shape[0] or cols != out.rows.shape[1]
                or (max(sizes) if sizes else 0) > out.max_gt):
            raise ValueError("ground truth does not fit the static buffers")
        out.offsets.copy_(off, non_blocking=True)
        if total:
            out.rows[:total].copy_(rows_host, non_blocking=True)
        _devcache.mark_in_flight()
        return out
    dev_words = stage[: head + total * cols].to(device, non_blocking=True)
    _devcache.mark_in_flight()
    offsets_dev = dev_words[: batch + 1].view(torch.int32)
    rows_dev = dev_words[head:].view(total, cols)
    return PackedGroundTruth(rows_dev, offsets_dev, max(sizes) if sizes else 0, batch)


class TargetAssigner(object):
    """detection/target_assigner.py:17-63.

    ``nan_check`` controls the reference's runtime assert (no NaN in positive target boxes,
    target_assigner.py:60-61), which costs a host sync there: ``"sync"`` raises AssertionError
    before returning, ``"deferred"`` (default) raises at the next call or at :meth:`check`,
    ``"off"`` skips it.  The NaN count comes out of the kernel's statistics either way.
    """

    def __init__(self, matched_threshold, unmatched_threshold, nan_check: str = "deferred"):
        self.matched_threshold = matched_threshold
        self.unmatched_threshold = unmatched_threshold
        self.nan_check = nan_check
        self.last_match: Optional[torch.Tensor] = None      # int32 [B, A] matcher output
        self.last_stats: Optional[torch.Tensor] = None      # int32 [B, 4] positives, ignored, nan, G
        self._pending = None

    def check(self) -> None:
        if self._pending is not None:
            stats_host, event = self._pending
            self._pending = None
            event.synchronize()
            assert int(stats_host[:, 2].sum()) == 0, "NaN in the target box of a positive anchor"

    def encode_packed(self, packed: PackedGroundTruth, anchors: torch.Tensor, box_coder=None) -> torch.Tensor:
        """Device-resident entry point: no host packing, no sync.

        ``box_coder``: write the box columns already coded for the loss -- what
        ``to_centroids(target_locs, inplace=True); box_coder.encode_box(target_locs, anchors, inplace=True)``
        (multibox_loss.py:81-82) would turn them into, bit for bit -- in the same launch."""
        assert self.matched_threshold >= self.unmatched_threshold           # matcher.py:43
        coding = None if box_coder is None else [float(box_coder.xy_scale), float(box_coder.wh_scale), float(box_coder.eps)]
        target, match, stats = OPS.assign_targets(anchors, packed.rows, packed.offsets, packed.max_gt,
                                                  float(self.matched_threshold), float(self.unmatched_threshold), True,
                                                  coding)
        self.last_match, self.last_stats = match, stats
        return target

    def encode_ground_truth(self, ground_truth, anchors):
        """
        Args:
            ground_truth: list(:len Batch) of torch.tensor(:shape [Boxes_i, >=6])
            anchors: torch.tensor(:shape [AnchorBoxes, 4])
        Returns:
            target: torch.tensor(:shape [Batch, AnchorBoxes, 6]) on the GPU
        """
        self.check()
        device = anchors.device if anchors.is_cuda else torch.device("cuda", torch.cuda.current_device())
        anchors_dev = _devcache.device_copy(anchors, device)
        packed = pack_ground_truth(ground_truth, device)
        target = self.encode_packed(packed, anchors_dev)
        if self.nan_check != "off":
            stats_host = _devcache.pinned_buffer("assign_stats", self.last_stats.shape, torch.int32)
            stats_host.copy_(self.last_stats, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            self._pending = (stats_host, event)
            if self.nan_check == "sync":
                self.check()
        return target
