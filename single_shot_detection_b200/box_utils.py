"""Box utilities -- same interface as the reference's ``bf/utils/box_utils.py`` for CUDA tensors.

``to_corners`` / ``to_centroids`` / ``iou`` / ``nms`` run hand-written kernels through the C ABI;
``area`` and ``intersection`` are one-liners of element-wise torch ops kept for API completeness
(nothing on the batch path calls them).  CPU tensors and numpy arrays are rejected: this package
has no CPU route, the reference's own module keeps serving the data-loader augmentations
(``bf/preprocessing/functional/box.py:68-69``).
"""
import torch

from . import _native as N
from .ops import OPS


def _need_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"single_shot_detection_b200.box_utils.{name} needs CUDA tensors (no CPU fallback); "
                        "use the reference's bf.utils.box_utils for numpy / CPU inputs")


def to_corners(box):
    """(cx,cy,w,h) -> (x1,y1,x2,y2).  bf/utils/box_utils.py:16-23"""
    _need_cuda(box, "to_corners")
    return OPS.box_transform(box, None, N.BOX_TO_CORNERS, 1.0, 1.0, 0.0)


def to_centroids(box, inplace=False):
    """(x1,y1,x2,y2) -> (cx,cy,w,h).  bf/utils/box_utils.py:25-36 (both rounding orders)."""
    _need_cuda(box, "to_centroids")
    if inplace:
        OPS.box_transform_(box, None, N.BOX_TO_CENTROIDS_INPLACE, 1.0, 1.0, 0.0)
        return None
    return OPS.box_transform(box, None, N.BOX_TO_CENTROIDS, 1.0, 1.0, 0.0)


def area(box):
    """bf/utils/box_utils.py:38-46"""
    return (box[..., 2] - box[..., 0]).clamp_(0) * (box[..., 3] - box[..., 1]).clamp_(0)


def intersection(a, b, cartesian=True, zero_incorrect=False):
    """Intersection rectangles, bf/utils/box_utils.py:49-80."""
    _need_cuda(a, "intersection")
    if cartesian:
        lo = torch.maximum(a[:, None, :2], b[None, :, :2])
        hi = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    else:
        assert a.size() == b.size()
        lo = torch.maximum(a[..., :2], b[..., :2])
        hi = torch.minimum(a[..., 2:], b[..., 2:])
    out = torch.cat([lo, hi], dim=-1)
    if zero_incorrect:
        out[(hi < lo).any(dim=-1)] = 0
    return out


def iou(a, b, cartesian=True):
    """IoU of corner boxes, bf/utils/box_utils.py:83-101.  [BoxesA, BoxesB] when cartesian."""
    _need_cuda(a, "iou")
    if cartesian:
        return OPS.pairwise_iou(a, b)
    inter = area(intersection(a, b, cartesian=False))
    return inter / (area(a) + area(b) - inter)


def nms(boxes, scores, overlap_threshold, score_threshold, max_per_class=None, soft=False, sigma=0.5):
    """Top-k + hard NMS for one box set, bf/utils/box_utils.py:165-194.

    Returns ((boxes_picked, scores_picked), indexes_picked); ``indexes_picked`` index the INPUT rows
    (the reference indexes its unsorted top-k subset, whose order is implementation defined).
    """
    _need_cuda(boxes, "nms")
    if soft:
        raise NotImplementedError("soft-NMS (box_utils.py:145-163) is not part of the accelerated path")
    n = int(scores.shape[0])
    k = n if max_per_class is None else min(int(max_per_class), n)
    if n == 0:
        empty = torch.zeros((0,), dtype=torch.long, device=boxes.device)
        return (boxes[empty], scores[empty]), empty
    keep, count = OPS.nms(boxes, scores, max(k, 1), float(overlap_threshold))
    picked = keep[: int(count.item())]
    return (boxes[picked], scores[picked]), picked
