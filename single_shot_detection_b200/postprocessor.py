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
        if nms.get('max_per_class') is None:
            raise NotImplementedError('max_per_class=None (NMS over every candidate) is not supported; '
                                      'every reference sample sets it (100)')
        self._nms_cfg = dict(nms)
        self.last_status = None

    def postprocess_padded(self, prediction, priors):
        """No-sync variant: (dets [B, cap, 6], counts [B] int32, anchors [B, cap] int32, status [4])."""
        b_scores, b_boxes = prediction
        device = b_scores.device
        if not b_scores.is_cuda:
            raise TypeError('Postprocessor needs CUDA predictions (no CPU fallback)')
        priors_dev = _devcache.device_copy(priors, device)
        converter, first_fg = _CONVERTERS[self.score_converter]
        max_total = int(self.max_total) if self.max_total is not None else 0
        return OPS.postprocess(b_scores, b_boxes, priors_dev, converter, first_fg, N.BOXES_ENCODED,
                               float(self.box_coder.xy_scale), float(self.box_coder.wh_scale),
                               float(self.score_threshold), int(self._nms_cfg['max_per_class']),
                               float(self._nms_cfg['overlap_threshold']), max_total,
                               float(self._nms_cfg.get('sigma', 0.5)) if self._nms_cfg.get('soft', False) else 0.0)

    def begin_padded(self, prediction, priors, want_loss_keys=False):
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
        return self.to_list(*self.postprocess_padded(prediction, priors))

    def to_list(self, dets, counts, anchors, status):
        """Padded device output -> the reference's list of ``[n_i, 6]`` views (one host sync: the counts)."""
        self.last_padded = (dets, counts, anchors, status)
        host = _devcache.pinned_buffer("post_counts", (counts.numel() + 4,), torch.int32)
        host[: counts.numel()].copy_(counts, non_blocking=True)
        host[counts.numel():].copy_(status, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.last_status = host[counts.numel():].tolist()
        if self.last_status[0]:
            raise RuntimeError(f'postprocess: {self.last_status[0]} (image, class) candidate lists overflowed; '
                               'result would be inexact')
        self.last_anchors = anchors
        return [dets[i, :n] for i, n in enumerate(host[: counts.numel()].tolist())]
