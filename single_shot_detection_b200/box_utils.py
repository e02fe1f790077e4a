"""Box utilities -- same interface as the reference's ``bf/utils/box_utils.py`` for CUDA tensors.

``to_corners`` / ``to_centroids`` / ``iou`` / ``generalized_iou`` / ``nms`` (hard and soft) run hand-written kernels through the C ABI;
``area`` and ``intersection`` are one-liners of element-wise torch ops kept for API completeness
(nothing on the batch path calls them).  CPU tensors are rejected: this package has no CPU route.
numpy arrays take the reference's ``@to_torch`` route (``intersection`` / ``iou`` / ``nms``): staged to
the GPU, computed by the same kernels, returned as numpy.
"""
import torch

from . import _native as N
from .ops import OPS


def to_torch(func):
    """bf/utils/box_utils.py:8-14: numpy in -> numpy out (the crop augmentation's route,
    bf/preprocessing/functional/box.py:68-69).  The arrays are staged onto the current CUDA device,
    the same kernels run, and the result comes back in the input's dtype -- there is still no CPU
    arithmetic here.  (Inside forked DataLoader workers CUDA is unusable: keep the reference's module
    for that process.)"""
    import functools

    import numpy as np

    @functools.wraps(func)
    def wrapped_function(*args, **kwargs):
        if isinstance(args[0], np.ndarray):
            dev = torch.device("cuda", torch.cuda.current_device())
            out = func(*[torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in args], **kwargs)

            def back(o):
                if isinstance(o, tuple):
                    return tuple(back(x) for x in o)
                o = o.cpu().numpy()
                return o.astype(args[0].dtype) if o.dtype.kind == "f" else o
            return back(out)
        return func(*args, **kwargs)
    return wrapped_function


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


@to_torch
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


@to_torch
def iou(a, b, cartesian=True):
    """IoU of corner boxes, bf/utils/box_utils.py:83-101.  [BoxesA, BoxesB] when cartesian."""
    _need_cuda(a, "iou")
    if cartesian:
        return OPS.pairwise_iou(a, b)
    inter = area(intersection(a, b, cartesian=False))
    return inter / (area(a) + area(b) - inter)


def generalized_iou(a, b, cartesian=True):
    """GIoU of corner boxes, bf/utils/box_utils.py:104-143 (forward value; no autograd)."""
    _need_cuda(a, "generalized_iou")
    assert a.dim() == b.dim() == 2
    assert a.size(1) == b.size(1) == 4
    if not cartesian:
        assert a.size() == b.size()
    return OPS.generalized_iou(a, b, bool(cartesian))


@to_torch
def nms(boxes, scores, overlap_threshold, score_threshold, max_per_class=None, soft=False, sigma=0.5):
    """Top-k + hard NMS (or Gaussian soft-NMS, ``soft=True``) for one box set, bf/utils/box_utils.py:145-194.

    Returns ((boxes_picked, scores_picked), indexes_picked); ``indexes_picked`` index the INPUT rows
    (the reference indexes its unsorted top-k subset, whose order is implementation defined).
    """
    _need_cuda(boxes, "nms")
    n = int(scores.shape[0])
    k = n if max_per_class is None else min(int(max_per_class), n)
    if n == 0:
        empty = torch.zeros((0,), dtype=torch.long, device=boxes.device)
        return (boxes[empty], scores[empty]), empty
    if k > N.MAX_PER_CLASS:
        # more boxes enter the NMS than the batched kernel keeps in shared memory (max_per_class=None on a long list,
        # or a large max_per_class): the global-memory route, same semantics (csrc/nms_large.cu)
        keep, count = OPS.nms_large(boxes, scores, 0 if max_per_class is None else int(max_per_class),
                                    float(overlap_threshold), bool(soft), float(score_threshold), float(sigma))
    elif soft:            # box_utils.py:145-163 (picked in pick order; the returned scores are the INPUT scores)
        keep, count = OPS.soft_nms(boxes, scores, max(k, 1), float(score_threshold), float(sigma))
    else:
        keep, count = OPS.nms(boxes, scores, max(k, 1), float(overlap_threshold))
    picked = keep[: int(count.item())]
    return (boxes[picked], scores[picked]), picked
