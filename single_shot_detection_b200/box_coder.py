"""BoxCoder -- same interface as the reference's ``detection/box_coder.py``.

Both rounding orders of the reference exist in the kernel (csrc/boxes.cu): the in-place branch
``log(wh / p_wh + eps)`` used by the loss (multibox_loss.py:81-82) and the out-of-place branch
``log((wh + eps) / p_wh)``.  In-place calls work on strided views such as ``target[..., 0:4]``
(row stride 6) without a copy.
"""
import torch

from . import _devcache
from . import _native as N
from .ops import OPS


class BoxCoder(torch.nn.Module):
    __constants__ = ['xy_scale', 'wh_scale']

    def __init__(self, xy_scale, wh_scale, eps=1e-8):
        super(BoxCoder, self).__init__()
        self.xy_scale = xy_scale
        self.wh_scale = wh_scale
        self.eps = eps

    @staticmethod
    def _priors_for(boxes, priors):
        return _devcache.device_copy(priors, boxes.device)

    def encode_box(self, boxes, priors, inplace=False):
        """
        Args:
            boxes: torch.tensor(:shape [Batch, AnchorBoxes, 4]) centroids
            priors: torch.tensor(:shape [AnchorBoxes, 4])
        Returns:
            encoded: torch.tensor(:shape [Batch, AnchorBoxes, 4])
        """
        priors = self._priors_for(boxes, priors)
        if inplace:
            OPS.box_transform_(boxes, priors, N.BOX_ENCODE_INPLACE, float(self.xy_scale), float(self.wh_scale),
                               float(self.eps))
            return boxes
        return OPS.box_transform(boxes, priors, N.BOX_ENCODE, float(self.xy_scale), float(self.wh_scale),
                                 float(self.eps))

    def decode_box(self, boxes, priors, inplace=torch.tensor(0)):
        """
        Args:
            boxes: torch.tensor(:shape [Batch, AnchorBoxes, 4]) encoded locs
            priors: torch.tensor(:shape [AnchorBoxes, 4])
        Returns:
            decoded: torch.tensor(:shape [Batch, AnchorBoxes, 4]) centroids
        """
        priors = self._priors_for(boxes, priors)
        if inplace:
            OPS.box_transform_(boxes, priors, N.BOX_DECODE_INPLACE, float(self.xy_scale), float(self.wh_scale),
                               float(self.eps))
            return boxes
        return OPS.box_transform(boxes, priors, N.BOX_DECODE, float(self.xy_scale), float(self.wh_scale),
                                 float(self.eps))

    def encode_corners_(self, corner_boxes, priors):
        """Fused ``box_utils.to_centroids(x, inplace=True); encode_box(x, priors, inplace=True)`` --
        the two calls multibox_loss.py:81-82 makes -- in one pass over the rows (same rounding)."""
        priors = self._priors_for(corner_boxes, priors)
        OPS.box_transform_(corner_boxes, priors, N.BOX_CENTROIDS_ENCODE_INPLACE, float(self.xy_scale),
                           float(self.wh_scale), float(self.eps))
        return corner_boxes
