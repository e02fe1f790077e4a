"""Matcher -- same interface as the reference's ``detection/matcher.py``.

The batch path does not come through here (``TargetAssigner`` fuses IoU + matching in one
launch); these functions exist so that code calling the matcher directly keeps working on GPU
tensors.
"""
import torch

from .ops import OPS

NOT_MATCHED = -2
IGNORE = -1


def match_bipartite(weights, inplace=False):
    """Greedy bipartite matching, detection/matcher.py:7-31 (no caller in the reference).

    Args:
        weights: torch.tensor(:shape [Boxes, AnchorBoxes])
    Returns:
        box_idx: torch.tensor(:shape [Boxes])
        anchor_idx: torch.tensor(:shape [Boxes])
    """
    assert weights.max(dim=1)[0].gt(0).all().item()
    if not inplace:
        weights = weights.clone()
    num_boxes, num_priors = weights.size()
    box_idx = torch.arange(num_boxes, dtype=torch.long, device=weights.device)
    anchor_idx = torch.empty((num_boxes,), dtype=torch.long, device=weights.device)
    for _ in range(num_boxes):
        flat = weights.argmax()
        row, col = flat // num_priors, flat % num_priors
        anchor_idx[row] = col
        weights[:, col] = 0
        weights[row] = 0
    return box_idx, anchor_idx


def match_per_prediction(weights, matched_threshold, unmatched_threshold=None, force_match_for_each_target=True):
    """detection/matcher.py:33-56 on a CUDA ``weights[Boxes, AnchorBoxes]``.

    Returns:
        box_idx: torch.tensor(:shape [AnchorBoxes]) int64 in {-2, -1, 0..Boxes-1}
    """
    if unmatched_threshold is None:
        unmatched_threshold = matched_threshold
    else:
        assert matched_threshold >= unmatched_threshold
    return OPS.match_per_prediction(weights, float(matched_threshold), float(unmatched_threshold),
                                    bool(force_match_for_each_target))
