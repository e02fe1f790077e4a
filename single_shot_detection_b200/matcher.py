"""Matcher -- the interface of the reference's ``detection/matcher.py`` on GPU tensors.

The batch path does not come through here (``TargetAssigner`` fuses IoU + matching in one launch,
csrc/assign.cu); these two entry points exist so that code calling the matcher directly keeps working.
"""
import torch

from .ops import OPS

NOT_MATCHED = -2          # detection/matcher.py:4
IGNORE = -1               # detection/matcher.py:5


def match_bipartite(weights, inplace=False):
    """Greedy one-to-one matching (detection/matcher.py:7-31; the reference itself never calls it):
    ``Boxes`` rounds of "take the globally largest remaining weight, retire its row and its column".

    ``weights`` [Boxes, AnchorBoxes] -> (box_idx [Boxes] = 0..Boxes-1, anchor_idx [Boxes]), both int64.
    A retired entry counts as weight 0, exactly as the reference zeroes it, so once only non-positive weights
    are left the arg-max may land on a retired entry again -- the same caveat as there (its assert only
    guarantees one positive weight per box at the start).  The loop stays on the device; there is no host
    round trip inside it.
    """
    rows, cols = weights.shape
    if not bool((weights.amax(dim=1) > 0).all()):                       # matcher.py:15
        raise AssertionError("every box needs at least one positive weight")
    work = weights if inplace else weights.clone()
    picked = weights.new_empty((rows,), dtype=torch.long)
    zero = work.new_zeros(())
    for _round in range(rows):
        best = torch.argmax(work)                                       # first maximum in row-major order
        r, c = torch.div(best, cols, rounding_mode="floor"), torch.remainder(best, cols)
        picked.index_copy_(0, r.view(1), c.view(1))
        work.index_fill_(1, c.view(1), zero)
        work.index_fill_(0, r.view(1), zero)
    return torch.arange(rows, dtype=torch.long, device=weights.device), picked


def match_per_prediction(weights, matched_threshold, unmatched_threshold=None, force_match_for_each_target=True):
    """Best box per anchor with the two IoU thresholds and the forced match of every box to its best anchor
    (detection/matcher.py:33-56), one launch: ``weights`` [Boxes, AnchorBoxes] on the GPU -> int64
    [AnchorBoxes] holding a box index, ``IGNORE`` or ``NOT_MATCHED``.  A single threshold means no ignore band.
    """
    low = matched_threshold if unmatched_threshold is None else unmatched_threshold
    assert matched_threshold >= low                                      # matcher.py:43
    return OPS.match_per_prediction(weights, float(matched_threshold), float(low), bool(force_match_for_each_target))
