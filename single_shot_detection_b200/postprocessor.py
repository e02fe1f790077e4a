"""Postprocessor -- same interface as the reference's ``detection/postprocessor.py``.

``postprocess((scores, locs), priors)`` returns the reference's ``list[Tensor[n_i, 6]]``
(x1,y1,x2,y2,class,score) -- views of one padded device buffer.  The B x (C-1) Python loop of the
reference (threshold -> boolean gather -> top-k -> torchvision nms per class) is replaced by five
launches for the whole batch (csrc/postprocess.cu); the only host sync is the read of the B row
counts needed to cut the list.  ``postprocess_padded`` skips even that.
"""
import functools

import torch

from . import _devcache
from . import _native as N
from . import box_utils
from .ops import OPS

_CONVERTERS = {'SOFTMAX': (N.CONVERT_SOFTMAX, 1), 'SIGMOID': (N.CONVERT_SIGMOID, 0)}


class Postprocessor(object):
    def __init__(self, box_coder, score_threshold, nms, score_converter='SOFTMAX', max_total=None):
        self.box_coder = box_coder
        self.score_threshold = score_threshold
        self.nms = functools.partial(box_utils.nms, score_threshold=score_threshold, **nms)
        self.max_total = max_total
        self.score_converter = score_converter
        if score_converter not in _CONVERTERS:
            raise ValueError(f'Wrong value for score_converter: {score_converter}')
        self._nms_cfg = dict(nms)
        # The batched kernels keep max_per_class <= 512 boxes per (image, class) in shared memory.  max_per_class=None
        # (NMS over EVERY box above the threshold) or a larger bound -- no reference sample uses either -- take the
        # reference's own loop structure over (image, class) with the long-list NMS (csrc/nms_large.cu) inside.
        k = nms.get('max_per_class')
        self._unbounded = k is None or int(k) > N.MAX_PER_CLASS
        self.last_status = None

    def postprocess_padded(self, prediction, priors):
        """No-sync variant: (dets [B, cap, 6], counts [B] int32, anchors [B, cap] int32, status [4])."""
        b_scores, b_boxes = prediction
        device = b_scores.device
        if not b_scores.is_cuda:
            raise TypeError('Postprocessor needs CUDA predictions (no CPU fallback)')
        if self._unbounded:
            raise NotImplementedError('postprocess_padded needs max_per_class <= %d; postprocess() handles the rest'
                                      % N.MAX_PER_CLASS)
        priors_dev = _devcache.device_copy(priors, device)
        converter, first_fg = _CONVERTERS[self.score_converter]
        max_total = int(self.max_total) if self.max_total is not None else 0
        return OPS.postprocess(b_scores, b_boxes, priors_dev, converter, first_fg, N.BOXES_ENCODED,
                               float(self.box_coder.xy_scale), float(self.box_coder.wh_scale),
                               float(self.score_threshold), int(self._nms_cfg['max_per_class']),
                               float(self._nms_cfg['overlap_threshold']), max_total,
                               float(self._nms_cfg.get('sigma', 0.5)) if self._nms_cfg.get('soft', False) else 0.0)

    def begin_padded(self, prediction, priors, want_loss_keys=False):
        if self._unbounded:
            raise NotImplementedError('begin_padded needs max_per_class <= %d' % N.MAX_PER_CLASS)
        return self._begin_padded(prediction, priors, want_loss_keys)

    def _begin_padded(self, prediction, priors, want_loss_keys=False):
        """First launch only (row statistics; with ``want_loss_keys`` also the sampler's criterion, see
        ops.postprocess_begin); ``.finish()`` on the result enqueues the rest and returns what
        :meth:`postprocess_padded` returns."""
        from . import ops
        b_scores, b_boxes = prediction
        if not b_scores.is_cuda:
            raise TypeError('Postprocessor needs CUDA predictions (no CPU fallback)')
        priors_dev = _devcache.device_copy(priors, b_scores.device)
        converter, first_fg = _CONVERTERS[self.score_converter]
        max_total = int(self.max_total) if self.max_total is not None else 0
        return ops.postprocess_begin(b_scores, b_boxes, priors_dev, converter, first_fg, N.BOXES_ENCODED,
                                     float(self.box_coder.xy_scale), float(self.box_coder.wh_scale),
                                     float(self.score_threshold), int(self._nms_cfg['max_per_class']),
                                     float(self._nms_cfg['overlap_threshold']), max_total,
                                     float(self._nms_cfg.get('sigma', 0.5)) if self._nms_cfg.get('soft', False) else 0.0,
                                     want_loss_keys)

    def postprocess(self, prediction, priors):
        """
        Args:
            prediction: tuple of
                torch.tensor(:shape [Batch, AnchorBoxes * Classes])
                torch.tensor(:shape [Batch, AnchorBoxes * 4])
            priors: torch.tensor(:shape [AnchorBoxes, 4]
        Returns:
            processed: list(:len Batch) of torch.tensor(:shape [Boxes_i, 6] ~ {[0-3] - box, [4] - class, [5] - score})
        """
        if self._unbounded:
            return self._postprocess_unbounded(prediction, priors)
        return self.to_list(*self.postprocess_padded(prediction, priors))

    def _postprocess_unbounded(self, prediction, priors):
        """detection/postprocessor.py:36-76 with an unbounded (or > 512) max_per_class: the conversion, decoding
        and the loop over (image, class) as the reference writes them (device tensor ops + this package's decode
        kernel), ``box_utils.nms`` -- the long-list kernels -- per class.  One host sync per class, like the
        reference's boolean indexing; this route exists for completeness, not speed."""
        b_scores, b_boxes = prediction
        if not b_scores.is_cuda:
            raise TypeError('Postprocessor needs CUDA predictions (no CPU fallback)')
        device = b_scores.device
        batch_size, num_priors = b_scores.size(0), priors.size(0)
        priors_dev = _devcache.device_copy(priors, device)
        b_scores = b_scores.float().view(batch_size, num_priors, -1)
        b_scores = torch.sigmoid(b_scores) if self.score_converter == 'SIGMOID' else torch.softmax(b_scores, dim=-1)
        if self.score_converter == 'SOFTMAX':
            b_scores = b_scores[..., 1:]
        num_classes = b_scores.size(-1)
        b_boxes = self.box_coder.decode_box(b_boxes.float().view(batch_size, num_priors, 4), priors_dev,
                                            inplace=torch.tensor(0))
        b_boxes = box_utils.to_corners(b_boxes)
        processed = []
        for scores, boxes in zip(b_scores, b_boxes):
            picked = []
            for class_index in range(num_classes):
                class_scores = scores[:, class_index].contiguous()
                mask = class_scores > self.score_threshold
                (boxes_picked, scores_picked), _ = self.nms(boxes[mask], class_scores[mask])
                classes_picked = torch.full_like(scores_picked.unsqueeze(1), class_index + 1, dtype=torch.float)
                picked.append(torch.cat([boxes_picked, classes_picked, scores_picked.unsqueeze(1)], dim=-1))
            picked = torch.cat(picked, dim=0)
            if self.max_total is not None and self.max_total < picked.size(0):
                _, indexes = torch.topk(picked[:, 5], self.max_total, sorted=True, largest=True)
                picked = picked[indexes]
            processed.append(picked)
        self.last_status = [0, 0, 0, 0]
        return processed

    def to_list(self, dets, counts, anchors, status):
        """Padded device output -> the reference's list of ``[n_i, 6]`` views (one host sync: the counts)."""
        self.last_padded = (dets, counts, anchors, status)
        host = _devcache.pinned_buffer("post_counts", (counts.numel() + 4,), torch.int32)
        host[: counts.numel()].copy_(counts, non_blocking=True)
        host[counts.numel():].copy_(status, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        # status words (ssd_postprocess): [0] unused (always 0), [1] (image, class) lists that overflowed and were
        # redone exactly from the score column, [2] lists too long for the rank sort (bitonic path), [3] longest list
        self.last_status = host[counts.numel():].tolist()
        self.last_anchors = anchors
        return [dets[i, :n] for i, n in enumerate(host[: counts.numel()].tolist())]
