"""MultiboxLoss -- same interface as the reference's ``detection/losses/multibox_loss.py``.

``forward(pred, anchors, target)`` returns ``(loss, class_loss, loc_loss)`` exactly as the
reference does (multibox_loss.py:35-94): the sampler picks the anchors of the classification term,
``to_centroids`` + ``encode_box`` turn the target boxes into regression targets IN PLACE (the
reference mutates ``target`` too, :81-82), both sums are scaled by their weight and divided by
the number of positives.

The reference gathers ``scores[sampled_mask]`` / ``locs[positive_mask]`` through PyTorch, runs
the loss modules and lets autograd scatter the gradients back.  Here one fused pass
(csrc/loss.cu, ``ssd_multibox_loss``) produces the three loss values AND the dense gradients
w.r.t. ``scores`` and ``locs``; ``torch.autograd`` only sees one custom Function.

Supported configurations: ``CrossEntropyLoss`` or ``SigmoidFocalLoss`` for classification (what the
samples use), ``SmoothL1Loss`` or ``GeneralizedIoULoss`` for localisation -- with the latter the
target boxes stay corner boxes and the predictions are decoded instead (multibox_loss.py:77-79), the
gradient flows through GIoU, ``to_corners`` and the decoding inside the same kernel.  The soft-target
classification losses raise: there is no silent fallback.
"""
import torch
import torch.nn as nn

from . import _native as N
from . import box_utils
from .ops import OPS
from .target_assigner import LOC_INDEX_START, LOC_INDEX_END, CLASS_INDEX  # noqa: F401


class _FusedMultiboxLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, locs, target, sampled_mask, kind, gamma, alpha, class_weight, loc_weight,
                giou_priors=None, xy_scale=1.0, wh_scale=1.0):
        need_grad = scores.requires_grad or locs.requires_grad
        loss3, grad_scores, grad_locs = OPS.multibox_loss(scores, locs, target, sampled_mask, kind, gamma, alpha,
                                                          class_weight, loc_weight, need_grad, giou_priors,
                                                          xy_scale, wh_scale)
        ctx.save_for_backward(grad_scores, grad_locs)
        ctx.dtypes = (scores.dtype, locs.dtype)
        return loss3

    @staticmethod
    def backward(ctx, g):
        grad_scores, grad_locs = ctx.saved_tensors
        # loss3 = (class + loc, class, loc): d/d scores flows through entries 0 and 1, d/d locs through 0 and 2
        gs = gl = None
        if ctx.needs_input_grad[0]:
            gs = (grad_scores * (g[0] + g[1])).to(ctx.dtypes[0])
        if ctx.needs_input_grad[1]:
            gl = (grad_locs * (g[0] + g[2])).to(ctx.dtypes[1])
        return gs, gl, None, None, None, None, None, None, None, None, None, None


class MultiboxLoss(nn.Module):
    def __init__(self,
                 sampler,
                 box_coder,
                 classification_loss,
                 localization_loss,
                 classification_weight=1.0,
                 localization_weight=1.0):
        super(MultiboxLoss, self).__init__()

        self.sampler = sampler
        self.box_coder = box_coder

        cls_cfg = dict(classification_loss)
        name = cls_cfg.pop('name')
        if name == 'CrossEntropyLoss':
            if cls_cfg:
                raise NotImplementedError(f'CrossEntropyLoss options {sorted(cls_cfg)} are not supported')
            self.kind, self.gamma, self.alpha = N.LOSS_SOFTMAX_CE, 0.0, 0.0
            self.multiclass = False
        elif name == 'SigmoidFocalLoss':
            self.kind = N.LOSS_SIGMOID_FOCAL
            self.gamma = float(cls_cfg.pop('gamma', 2.0))
            self.alpha = float(cls_cfg.pop('alpha', 0.25))
            if cls_cfg.pop('epsilon', 0.0) or cls_cfg:
                raise NotImplementedError('SigmoidFocalLoss: only gamma and alpha are supported')
            self.multiclass = True
        else:
            raise NotImplementedError(f'classification loss {name!r} is not part of the accelerated path')
        self.soft_target = False

        loc_cfg = dict(localization_loss)
        loc_name = loc_cfg.pop('name')
        if loc_name not in ('SmoothL1Loss', 'GeneralizedIoULoss') or loc_cfg:
            raise NotImplementedError(f'localization loss {localization_loss!r} is not part of the accelerated path')
        self.iou_loss = loc_name == 'GeneralizedIoULoss'               # bf/modules/losses.py:110 IOU_LOSS

        self.classification_weight = classification_weight
        self.localization_weight = localization_weight

    def forward(self, pred, anchors, target):
        """
        Args:
            pred: tuple of
                torch.tensor(:shape [Batch, AnchorBoxes * Classes])
                torch.tensor(:shape [Batch, AnchorBoxes * 4])
            target: torch.tensor(:shape [Batch, AnchorBoxes, 6])
        Returns:
            losses: tuple(float, float)
        """
        scores, locs = pred

        target_locs = target[..., LOC_INDEX_START:LOC_INDEX_END]
        target_classes = target[..., CLASS_INDEX].long()

        batch_size = target.size(0)
        num_priors = target.size(1)

        sampled_mask = self.sampler(scores.view(batch_size, num_priors, -1), target_classes)

        if self.iou_loss:
            # multibox_loss.py:77-79: the predictions are decoded (inside the kernel), the target stays as it is
            from . import _devcache
            priors = _devcache.device_copy(anchors, target.device)
            loss3 = _FusedMultiboxLoss.apply(scores, locs, target, sampled_mask, self.kind, self.gamma, self.alpha,
                                             float(self.classification_weight), float(self.localization_weight),
                                             priors, float(self.box_coder.xy_scale), float(self.box_coder.wh_scale))
            return loss3[0], loss3[1], loss3[2]

        box_utils.to_centroids(target_locs, inplace=True)
        self.box_coder.encode_box(target_locs, anchors, inplace=True)

        loss3 = _FusedMultiboxLoss.apply(scores, locs, target, sampled_mask, self.kind, self.gamma, self.alpha,
                                         float(self.classification_weight), float(self.localization_weight))
        return loss3[0], loss3[1], loss3[2]
