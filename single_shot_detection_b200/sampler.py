"""Samplers -- same interface as the reference's ``detection/sampler.py``.

``detection/init.py:90-92`` looks a sampler up by name and filters its config through
``sampler.__code__.co_varnames``, so these stay plain Python functions with the reference's
parameter names.
"""
import math

from .ops import OPS
from .target_assigner import NEGATIVE_CLASS, IGNORE_CLASS  # noqa: F401  (re-exported like the reference)


def naive_sampler(predictions, target_classes):
    return OPS.positive_mask(target_classes)


def hard_negative_mining(predictions, target_classes, negative_per_positive_ratio, min_negative_per_image):
    """3:1 online hard-negative mining, detection/sampler.py:12-25.

    predictions [B, A, C] logits, target_classes [B, A] int64 -> bool [B, A].  Loss values that tie
    exactly across the cut go to the lower anchor index (the reference's unstable argsort leaves
    that case implementation defined).

    Tolerance: the SELECTION is bit-exact on identical fp32 losses (``hard_negative_mining_from_loss``).  From
    logits, the criterion -log_softmax(x)[0] is evaluated with the hardware ex2 / lg2 approximations (relative
    error 2^-22, csrc/rowstream.cuh), so anchors whose losses differ by less than that can swap sides of the cut:
    at most a few anchors per image against the reference (counts per BASELINE configuration:
    profiles/r02_mining_mismatch.md; the tests bound it by 2 per image).
    """
    ratio_is_integer = isinstance(negative_per_positive_ratio, int) and not isinstance(negative_per_positive_ratio, bool)
    mask, stats = OPS.hard_negative_mask(predictions, target_classes, None, float(negative_per_positive_ratio),
                                         ratio_is_integer, float(min_negative_per_image))
    hard_negative_mining.last_stats = stats
    return mask


hard_negative_mining.last_stats = None


def hard_negative_mining_from_loss(loss, target_classes, negative_per_positive_ratio, min_negative_per_image):
    """Selection half only: ``loss`` [B, A] replaces -log_softmax(predictions)[..., 0].  Used to check
    the selection bit-exactly on identical fp32 inputs."""
    ratio_is_integer = isinstance(negative_per_positive_ratio, int)
    mask, _ = OPS.hard_negative_mask(None, target_classes, loss, float(negative_per_positive_ratio),
                                     ratio_is_integer, float(min_negative_per_image))
    return mask


def hard_negative_mining_from_keys(loss_keys, target, negative_per_positive_ratio, min_negative_per_image, match=None):
    """Selection half on the criterion the post-processor's first pass already produced from the same
    logits (``Postprocessor.begin_padded(..., want_loss_keys=True)``): in an eval step
    (detection/init.py:117-122) the loss and the post-processor see the same prediction, so the logits are
    streamed once for both.  ``target`` is the [B, A, 6] target (its class column is read in place) or the
    int64 class tensor; ``match`` (optional) the matcher output of the assignment that wrote ``target``."""
    from . import ops
    ratio_is_integer = isinstance(negative_per_positive_ratio, int) and not isinstance(negative_per_positive_ratio, bool)
    mask, stats = ops.hard_negative_mask_from_keys(loss_keys, target, float(negative_per_positive_ratio),
                                                   ratio_is_integer, float(min_negative_per_image), match)
    hard_negative_mining.last_stats = stats
    return mask
