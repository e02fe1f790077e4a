"""torch custom ops ``torch.ops.ssd_b200.*`` -- thin shims over the C ABI (include/ssd_b200.h).

Each op is registered for the CUDA dispatch key only: calling it with CPU tensors raises (there is
no CPU fallback).  An op takes the raw device pointers of its tensor arguments and the current
CUDA stream and makes exactly one C-ABI call; no op synchronises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _native as N

_LIB = torch.library.Library("ssd_b200", "DEF")

_LIB.define("pairwise_iou(Tensor a, Tensor b) -> Tensor")
_LIB.define("generalized_iou(Tensor a, Tensor b, bool cartesian) -> Tensor")
_LIB.define("match_per_prediction(Tensor weights, float matched_threshold, float unmatched_threshold, "
            "bool force_match) -> Tensor")
_LIB.define("assign_targets(Tensor anchors, Tensor gt_rows, Tensor gt_offsets, int max_gt, "
            "float matched_threshold, float unmatched_threshold, bool force_match, float[]? box_coding=None) "
            "-> (Tensor, Tensor, Tensor)")
_LIB.define("box_transform(Tensor src, Tensor? priors, int op, float xy_scale, float wh_scale, float eps) -> Tensor")
_LIB.define("box_transform_(Tensor(a!) boxes, Tensor? priors, int op, float xy_scale, float wh_scale, float eps) -> ()")
_LIB.define("positive_mask(Tensor target_classes) -> Tensor")
_LIB.define("hard_negative_mask(Tensor? logits, Tensor target_classes, Tensor? loss, float ratio, "
            "bool ratio_is_integer, float min_negatives) -> (Tensor, Tensor)")
_LIB.define("postprocess(Tensor scores, Tensor boxes, Tensor? priors, int converter, int first_fg_col, "
            "int box_input, float xy_scale, float wh_scale, float score_threshold, int max_per_class, "
            "float overlap_threshold, int max_total, float soft_sigma=0.0) -> (Tensor, Tensor, Tensor, Tensor)")
_LIB.define("nms(Tensor boxes, Tensor scores, int max_per_class, float overlap_threshold) -> (Tensor, Tensor)")
_LIB.define("soft_nms(Tensor boxes, Tensor scores, int max_per_class, float score_threshold, float sigma) -> (Tensor, Tensor)")
_LIB.define("nms_large(Tensor boxes, Tensor scores, int max_keep, float overlap_threshold, bool soft, float score_threshold, "
            "float sigma) -> (Tensor, Tensor)")
_LIB.define("multibox_loss(Tensor scores, Tensor locs, Tensor target, Tensor sampled_mask, int kind, float gamma, "
            "float alpha, float class_weight, float loc_weight, bool need_grad, Tensor? giou_priors=None, "
            "float xy_scale=1.0, float wh_scale=1.0) -> (Tensor, Tensor, Tensor)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_workspaces: Dict[Tuple[int, str], torch.Tensor] = {}
_slot = 0                    # see workspace_slot()


class workspace_slot:
    """``with workspace_slot(k):`` -- the scratch buffers handed out inside belong to slot ``k``.  Steps that
    may run CONCURRENTLY (two step graphs in flight on two streams) must use different slots; steps that
    serialise on one stream can share one."""

    def __init__(self, slot: int):
        self.slot, self.prev = int(slot), 0

    def __enter__(self):
        global _slot
        self.prev, _slot = _slot, self.slot
        return self

    def __exit__(self, *exc):
        global _slot
        _slot = self.prev
        return False
_retired: list = []          # outgrown buffers: CUDA graphs captured earlier still hold their addresses


def workspace(nbytes: int, device: torch.device, tag: str = "default") -> torch.Tensor:
    """A cached, 256-byte aligned scratch buffer on ``device`` (grown on demand, never shrunk).

    Growth is geometric and an outgrown buffer is kept alive: a step graph captured while it was current
    keeps replaying with its address (torch.cuda.graph() empties the allocator cache when a capture begins,
    which would unmap a freed one)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), f"{tag}#{_slot}" if _slot else tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        size = max(int(nbytes), 256)
        if buf is not None:
            size = max(size, 2 * buf.numel())
            _retired.append(buf)
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError(f"the '{tag}' workspace would have to grow to {size} bytes inside a CUDA graph capture; "
                               "run the step once eagerly with the largest shapes first")
        buf = torch.empty(size, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous (the reference calls .float() on its inputs, postprocessor.py:39-40)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _rows_view(t: torch.Tensor) -> Optional[Tuple[int, int]]:
    """(row_stride_in_floats, rows) if ``t[..., 4]`` is a uniformly strided set of 4-float rows."""
    if t.dtype != torch.float32 or t.shape[-1] != 4 or t.stride(-1) != 1:
        return None
    if t.dim() == 1:
        return 4, 1
    rs = t.stride(-2)
    expect = rs
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1 and t.stride(d) != expect:
            return None
        expect *= t.shape[d]
    rows = 1
    for d in range(t.dim() - 1):
        rows *= t.shape[d]
    if rs < 4 or rs % 2 or t.data_ptr() % 8:
        return None
    return rs, rows


# ---------------------------------------------------------------------------------------------
def _pairwise_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    N.require_device()
    a, b = _f32c(a), _f32c(b)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        N.check(N.lib().ssd_pairwise_iou(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream()))
    return out


def _generalized_iou(a: torch.Tensor, b: torch.Tensor, cartesian: bool) -> torch.Tensor:
    N.require_device()
    a, b = _f32c(a), _f32c(b)
    shape = (a.shape[0], b.shape[0]) if cartesian else (a.shape[0],)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        N.check(N.lib().ssd_generalized_iou(_ptr(a), a.shape[0], _ptr(b), b.shape[0], int(cartesian), _ptr(out),
                                            _stream()))
    return out


def _match_per_prediction(weights: torch.Tensor, matched_threshold: float, unmatched_threshold: float,
                          force_match: bool) -> torch.Tensor:
    N.require_device()
    w = _f32c(weights)
    out = torch.empty((w.shape[1],), dtype=torch.int64, device=w.device)
    with torch.cuda.device(w.device):
        N.check(N.lib().ssd_match_per_prediction(_ptr(w), w.shape[0], w.shape[1], matched_threshold,
                                                 unmatched_threshold, int(force_match), _ptr(out), _stream()))
    return out


def _assign_targets(anchors: torch.Tensor, gt_rows: torch.Tensor, gt_offsets: torch.Tensor, max_gt: int,
                    matched_threshold: float, unmatched_threshold: float, force_match: bool, box_coding=None):
    """``box_coding`` = (xy_scale, wh_scale, eps): the box columns come out already passed through the loss
    route's to_centroids + encode_box (ssd_assign_targets_encoded)."""
    N.require_device()
    anchors, gt_rows = _f32c(anchors), _f32c(gt_rows)
    if gt_offsets.dtype != torch.int32 or not gt_offsets.is_contiguous():
        gt_offsets = gt_offsets.to(torch.int32).contiguous()
    batch = gt_offsets.numel() - 1
    num_anchors = anchors.shape[0]
    dev = anchors.device
    target = torch.empty((batch, num_anchors, 6), dtype=torch.float32, device=dev)
    match = torch.empty((batch, num_anchors), dtype=torch.int32, device=dev)
    stats = torch.empty((batch, 4), dtype=torch.int32, device=dev)
    gt_cols = gt_rows.shape[1] if gt_rows.dim() == 2 else 6
    with torch.cuda.device(dev):
        cap = 64                                        # scratch sized for a bucket of boxes per image, not for
        while cap < max_gt:                             # this batch's maximum: its address stays put across batches
            cap *= 2
        ws_bytes = N.lib().ssd_assign_workspace_bytes(batch, cap)
        ws = workspace(ws_bytes, dev, "assign")
        if box_coding is not None:
            xy, wh, eps = (float(x) for x in box_coding)
            N.check(N.lib().ssd_assign_targets_encoded(_ptr(anchors), _ptr(gt_rows) if gt_rows.numel() else None, gt_cols,
                                                       _ptr(gt_offsets), max_gt, batch, num_anchors, matched_threshold,
                                                       unmatched_threshold, int(force_match), xy, wh, eps, _ptr(target),
                                                       _ptr(match), _ptr(stats), _ptr(ws), ws.numel(), _stream()))
        else:
            N.check(N.lib().ssd_assign_targets(_ptr(anchors), _ptr(gt_rows) if gt_rows.numel() else None, gt_cols,
                                               _ptr(gt_offsets), max_gt, batch, num_anchors, matched_threshold,
                                               unmatched_threshold, int(force_match), _ptr(target), _ptr(match),
                                               _ptr(stats), _ptr(ws), ws.numel(), _stream()))
    return target, match, stats


def _box_call(op: int, src: torch.Tensor, src_stride: int, dst: torch.Tensor, dst_stride: int,
              priors: Optional[torch.Tensor], rows: int, xy: float, wh: float, eps: float) -> None:
    num_anchors = priors.shape[0] if priors is not None else 1
    with torch.cuda.device(src.device):
        N.check(N.lib().ssd_box_transform(op, _ptr(src), src_stride, _ptr(dst), dst_stride, _ptr(priors), rows,
                                          num_anchors, xy, wh, eps, _stream()))


def _box_transform(src: torch.Tensor, priors: Optional[torch.Tensor], op: int, xy_scale: float, wh_scale: float,
                   eps: float) -> torch.Tensor:
    N.require_device()
    view = _rows_view(src)
    if view is None:
        src = _f32c(src)
        view = (4, src.numel() // 4)
    out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    _box_call(op, src, view[0], out, 4, priors, view[1], xy_scale, wh_scale, eps)
    return out


def _box_transform_(boxes: torch.Tensor, priors: Optional[torch.Tensor], op: int, xy_scale: float, wh_scale: float,
                    eps: float) -> None:
    N.require_device()
    view = _rows_view(boxes)
    if view is None:                      # exotic strides: round-trip through a packed copy
        tmp = boxes.float().contiguous()
        _box_call(op, tmp, 4, tmp, 4, priors, tmp.numel() // 4, xy_scale, wh_scale, eps)
        boxes.copy_(tmp)
        return
    _box_call(op, boxes, view[0], boxes, view[0], priors, view[1], xy_scale, wh_scale, eps)


def _positive_mask(target_classes: torch.Tensor) -> torch.Tensor:
    N.require_device()
    cls = target_classes if target_classes.dtype == torch.int64 else target_classes.long()
    cls = cls if cls.is_contiguous() else cls.contiguous()
    out = torch.empty(cls.shape, dtype=torch.bool, device=cls.device)
    with torch.cuda.device(cls.device):
        N.check(N.lib().ssd_positive_mask(_ptr(cls), cls.numel(), _ptr(out), _stream()))
    return out


def _hard_negative_mask(logits: Optional[torch.Tensor], target_classes: torch.Tensor, loss: Optional[torch.Tensor],
                        ratio: float, ratio_is_integer: bool, min_negatives: float):
    N.require_device()
    cls = target_classes if target_classes.dtype == torch.int64 else target_classes.long()
    cls = cls if cls.is_contiguous() else cls.contiguous()
    if cls.data_ptr() % 16:                      # the class ids ride the TMA ring: 16-byte aligned
        cls = cls.clone()
    batch, num_anchors = cls.shape
    dev = cls.device
    num_cols = 0
    if logits is not None:
        logits = _f32c(logits.detach())
        num_cols = logits.numel() // max(batch * num_anchors, 1)
    if loss is not None:
        loss = _f32c(loss.detach())
    mask = torch.empty((batch, num_anchors), dtype=torch.bool, device=dev)
    stats = torch.empty((batch, 4), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = N.lib().ssd_hard_negative_workspace_bytes(batch, num_anchors)
        ws = workspace(nbytes, dev, "mining")
        N.check(N.lib().ssd_hard_negative_mask(_ptr(logits), _ptr(cls), _ptr(loss), batch, num_anchors, num_cols,
                                               float(ratio), int(ratio_is_integer), float(min_negatives), _ptr(mask),
                                               _ptr(stats), _ptr(ws), ws.numel(), _stream()))
    return mask, stats


def det_capacity(num_fg: int, max_per_class: int, max_total: int) -> int:
    rows = num_fg * max_per_class
    return min(rows, max_total) if max_total > 0 else rows


def _post_params(scores, boxes, converter, first_fg_col, box_input, xy_scale, wh_scale, score_threshold, max_per_class,
                 overlap_threshold, max_total, soft_sigma):
    batch = scores.shape[0]
    num_anchors = boxes.numel() // max(4 * batch, 1)
    num_cols = scores.numel() // max(batch * num_anchors, 1)
    p = N.PostprocessParams()
    p.batch, p.num_anchors, p.num_cols = batch, num_anchors, num_cols
    p.converter, p.first_fg_col, p.box_input = converter, first_fg_col, box_input
    p.xy_scale, p.wh_scale, p.score_threshold = xy_scale, wh_scale, score_threshold
    p.max_per_class, p.overlap_threshold, p.max_total = max_per_class, overlap_threshold, max_total
    p.det_capacity = det_capacity(num_cols - first_fg_col, max_per_class, max_total)
    p.soft_nms, p.soft_sigma, p.soft_threshold = int(soft_sigma > 0.0), soft_sigma, score_threshold
    return p


def _post_workspace(p, dev):
    nbytes = N.lib().ssd_postprocess_workspace_bytes(ctypes.byref(p))
    if nbytes == 0:
        N.check(N.lib().ssd_postprocess(ctypes.byref(p), None, None, None, None, None, None, None, None, 0, None))
    return workspace(nbytes, dev, "postprocess")


def _post_finish(p, scores, boxes, priors, ws):
    dev = scores.device
    batch, cap = p.batch, p.det_capacity
    dets = torch.empty((batch, cap, 6), dtype=torch.float32, device=dev)
    counts = torch.empty((batch,), dtype=torch.int32, device=dev)
    anchors = torch.empty((batch, cap), dtype=torch.int32, device=dev)
    status = torch.empty((4,), dtype=torch.int32, device=dev)
    N.check(N.lib().ssd_postprocess(ctypes.byref(p), _ptr(scores), _ptr(boxes), _ptr(priors), _ptr(dets),
                                    _ptr(counts), _ptr(anchors), _ptr(status), _ptr(ws), ws.numel(), _stream()))
    return dets, counts, anchors, status


def _postprocess(scores: torch.Tensor, boxes: torch.Tensor, priors: Optional[torch.Tensor], converter: int,
                 first_fg_col: int, box_input: int, xy_scale: float, wh_scale: float, score_threshold: float,
                 max_per_class: int, overlap_threshold: float, max_total: int, soft_sigma: float = 0.0):
    N.require_device()
    scores = _f32c(scores.detach())
    boxes = _f32c(boxes.detach())
    p = _post_params(scores, boxes, converter, first_fg_col, box_input, xy_scale, wh_scale, score_threshold,
                     max_per_class, overlap_threshold, max_total, soft_sigma)
    with torch.cuda.device(scores.device):
        return _post_finish(p, scores, boxes, priors, _post_workspace(p, scores.device))


class PostprocessInFlight:
    """The post-processor split after its first launch (ssd_postprocess_pass1): ``loss_keys`` [B, A] int32
    holds the sampler's raw criterion keys when requested; :meth:`finish` enqueues the remaining launches
    on the CURRENT stream.  Between the two the caller may record an event, so that the selection of the
    sampler (another stream) waits for pass 1 only."""

    def __init__(self, p, scores, boxes, priors, ws, loss_keys):
        self.p, self.scores, self.boxes, self.priors, self.ws, self.loss_keys = p, scores, boxes, priors, ws, loss_keys

    def finish(self):
        self.p.resume_after_pass1 = 1
        with torch.cuda.device(self.scores.device):
            return _post_finish(self.p, self.scores, self.boxes, self.priors, self.ws)


def postprocess_begin(scores: torch.Tensor, boxes: torch.Tensor, priors: Optional[torch.Tensor], converter: int,
                      first_fg_col: int, box_input: int, xy_scale: float, wh_scale: float, score_threshold: float,
                      max_per_class: int, overlap_threshold: float, max_total: int, soft_sigma: float = 0.0,
                      want_loss_keys: bool = False) -> PostprocessInFlight:
    N.require_device()
    scores = _f32c(scores.detach())
    boxes = _f32c(boxes.detach())
    p = _post_params(scores, boxes, converter, first_fg_col, box_input, xy_scale, wh_scale, score_threshold,
                     max_per_class, overlap_threshold, max_total, soft_sigma)
    dev = scores.device
    keys = torch.empty((p.batch, p.num_anchors), dtype=torch.int32, device=dev) if want_loss_keys else None
    with torch.cuda.device(dev):
        ws = _post_workspace(p, dev)
        N.check(N.lib().ssd_postprocess_pass1(ctypes.byref(p), _ptr(scores), _ptr(keys), _ptr(ws), ws.numel(), _stream()))
    return PostprocessInFlight(p, scores, boxes, priors, ws, keys)


def hard_negative_mask_from_keys(loss_keys: torch.Tensor, target: torch.Tensor, ratio: float, ratio_is_integer: bool,
                                 min_negatives: float, match: Optional[torch.Tensor] = None):
    """Selection on the raw criterion keys of :func:`postprocess_begin`; the classes are read straight from
    the class column of ``target`` [B, A, 6] fp32 (no ``.long()`` copy) or from an int64 [B, A] tensor.  ``match``:
    the int32 [B, A] matcher output of the assignment that wrote ``target`` (TargetAssigner.last_match) -- unmatched
    and ignored anchors then take their class from it and only matched anchors touch the strided class column."""
    N.require_device()
    batch, num_anchors = loss_keys.shape
    dev = loss_keys.device
    if target.dtype == torch.int64:
        cls = target if target.is_contiguous() else target.contiguous()
        cls_ptr, stride = cls.data_ptr(), 0
    else:
        assert target.dtype == torch.float32 and target.is_contiguous() and target.shape[-1] >= 5
        cls_ptr, stride = target.data_ptr() + 4 * 4, int(target.shape[-1])
    mask = torch.empty((batch, num_anchors), dtype=torch.bool, device=dev)
    stats = torch.empty((batch, 4), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        if match is not None:
            assert match.dtype == torch.int32 and match.is_contiguous() and tuple(match.shape) == (batch, num_anchors)
        N.check(N.lib().ssd_hard_negative_mask_from_keys(loss_keys.data_ptr(), cls_ptr, stride, _ptr(match), batch, num_anchors,
                                                         float(ratio), int(ratio_is_integer), float(min_negatives),
                                                         _ptr(mask), _ptr(stats), _stream()))
    return mask, stats


def _nms(boxes: torch.Tensor, scores: torch.Tensor, max_per_class: int, overlap_threshold: float):
    N.require_device()
    boxes, scores = _f32c(boxes), _f32c(scores)
    n = scores.numel()
    dev = boxes.device
    keep = torch.empty((max_per_class,), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = N.lib().ssd_nms_workspace_bytes(n, max_per_class)
        ws = workspace(max(nbytes, 256), dev, "nms")
        N.check(N.lib().ssd_nms(_ptr(boxes), _ptr(scores), n, max_per_class, overlap_threshold, _ptr(keep),
                                _ptr(count), _ptr(ws), ws.numel(), _stream()))
    return keep, count


def _soft_nms(boxes: torch.Tensor, scores: torch.Tensor, max_per_class: int, score_threshold: float, sigma: float):
    N.require_device()
    boxes, scores = _f32c(boxes), _f32c(scores)
    n = scores.numel()
    dev = boxes.device
    keep = torch.empty((max_per_class,), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = N.lib().ssd_nms_workspace_bytes(n, max_per_class)
        ws = workspace(max(nbytes, 256), dev, "nms")
        N.check(N.lib().ssd_soft_nms(_ptr(boxes), _ptr(scores), n, max_per_class, score_threshold, sigma, _ptr(keep),
                                     _ptr(count), _ptr(ws), ws.numel(), _stream()))
    return keep, count


def _nms_large(boxes: torch.Tensor, scores: torch.Tensor, max_keep: int, overlap_threshold: float, soft: bool,
               score_threshold: float, sigma: float):
    """box_utils.nms over a long list (ssd_nms_large): keep [k] int64 input rows + count [1] int32."""
    N.require_device()
    boxes, scores = _f32c(boxes), _f32c(scores)
    n = scores.numel()
    dev = boxes.device
    k = max_keep if 0 < max_keep < n else n
    keep = torch.empty((max(k, 1),), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = N.lib().ssd_nms_large_workspace_bytes(n, max_keep)
        ws = workspace(max(nbytes, 256), dev, "nms_large")
        N.check(N.lib().ssd_nms_large(_ptr(boxes), _ptr(scores), n, max_keep, overlap_threshold, int(soft), score_threshold,
                                      sigma, _ptr(keep), _ptr(count), _ptr(ws), ws.numel(), _stream()))
    return keep, count


def _multibox_loss(scores: torch.Tensor, locs: torch.Tensor, target: torch.Tensor, sampled_mask: torch.Tensor,
                   kind: int, gamma: float, alpha: float, class_weight: float, loc_weight: float, need_grad: bool,
                   giou_priors: Optional[torch.Tensor] = None, xy_scale: float = 1.0, wh_scale: float = 1.0):
    """(loss3 [3], grad_scores like scores, grad_locs like locs); the gradients are empty tensors when
    ``need_grad`` is false.  ``target`` rows hold the ENCODED boxes (after to_centroids + encode_box)."""
    N.require_device()
    batch, num_anchors = int(target.shape[0]), int(target.shape[1])
    scores_c, locs_c = _f32c(scores.detach()), _f32c(locs.detach())
    target_c = _f32c(target.detach())
    mask = sampled_mask if sampled_mask.is_contiguous() else sampled_mask.contiguous()
    if mask.dtype != torch.bool and mask.dtype != torch.uint8:
        mask = mask != 0
    num_cols = scores_c.numel() // max(batch * num_anchors, 1)
    dev = target_c.device
    loss3 = torch.empty((3,), dtype=torch.float32, device=dev)
    grad_scores = torch.empty(scores.shape if need_grad else (0,), dtype=torch.float32, device=dev)
    grad_locs = torch.empty(locs.shape if need_grad else (0,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nbytes = N.lib().ssd_multibox_loss_workspace_bytes(batch, num_anchors)
        ws = workspace(nbytes, dev, "loss")
        if giou_priors is not None:
            priors_c = _f32c(giou_priors)
            N.check(N.lib().ssd_multibox_loss_giou(_ptr(scores_c), _ptr(locs_c), _ptr(target_c), _ptr(priors_c), _ptr(mask),
                                                   batch, num_anchors, num_cols, kind, gamma, alpha, class_weight,
                                                   loc_weight, xy_scale, wh_scale,
                                                   _ptr(grad_scores) if need_grad else None,
                                                   _ptr(grad_locs) if need_grad else None, _ptr(loss3), _ptr(ws),
                                                   ws.numel(), _stream()))
        else:
            N.check(N.lib().ssd_multibox_loss(_ptr(scores_c), _ptr(locs_c), _ptr(target_c), _ptr(mask), batch, num_anchors,
                                              num_cols, kind, gamma, alpha, class_weight, loc_weight,
                                              _ptr(grad_scores) if need_grad else None,
                                              _ptr(grad_locs) if need_grad else None, _ptr(loss3), _ptr(ws), ws.numel(),
                                              _stream()))
    return loss3, grad_scores, grad_locs


for _name, _fn in [("multibox_loss", _multibox_loss), ("pairwise_iou", _pairwise_iou), ("generalized_iou", _generalized_iou), ("match_per_prediction", _match_per_prediction),
                   ("assign_targets", _assign_targets), ("box_transform", _box_transform),
                   ("box_transform_", _box_transform_), ("positive_mask", _positive_mask),
                   ("hard_negative_mask", _hard_negative_mask), ("postprocess", _postprocess), ("nms", _nms), ("soft_nms", _soft_nms),
                   ("nms_large", _nms_large)]:
    _LIB.impl(_name, _fn, "CUDA")

OPS = torch.ops.ssd_b200
