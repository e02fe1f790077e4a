"""Anchor generators -- same classes and constructor arguments as the reference's
``detection/anchor_generators/ssd.py`` / ``retina_net.py`` -- writing the table on the GPU.

The reference builds ``[H, W, boxes, 4]`` per level on the CPU and ``Detector.generate_anchors``
(detection/detector.py:82-86) flattens + concatenates the levels on every forward pass; the anchor
pipeline then copies the table to the device.  Here :func:`generate_anchors` writes the whole
``[A, 4]`` table of all levels with ONE launch (csrc/anchors.cu), bit-identical to the CPU table
(tests/test_gpu_parity.py against tables the reference's own generators produced), cached per
(image size, feature-map sizes).  The handful of per-level (w, h) scalars are computed on the host
with the very operations the reference uses (fp32 tensor x Python double, ``math.sqrt`` in double).
"""
from __future__ import annotations

import logging
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N


class _AnchorGenerator(object):
    """detection/anchor_generators/_anchor_generator.py."""

    num_boxes: int

    def _level(self, img_size, feature_map_size) -> N.AnchorLevel:
        raise NotImplementedError

    def _generate_anchors(self, img_size, feature_map_size, device=None):
        layer_w, layer_h = feature_map_size
        table = generate_anchors([self], img_size, [feature_map_size], device)
        return table.view(layer_h, layer_w, self.num_boxes, 4)

    def generate(self, img, feature_map):
        """
        Args:
            img: torch.tensor(:shape [Batch, Channels, Height, Width])
            feature_map: torch.tensor(:shape [Batch, Channels, Height, Width])
        Returns:
            priors: torch.tensor(:shape [Height, Width, AspectRatios, 4]) on the GPU
        """
        img_size = img.size(3), img.size(2)
        feature_map_size = feature_map.size(3), feature_map.size(2)
        device = feature_map.device if feature_map.is_cuda else None
        return self._generate_anchors(img_size, feature_map_size, device)


def _fill_level(level: N.AnchorLevel, img_size, feature_map_size, step, offset, hws: torch.Tensor) -> N.AnchorLevel:
    img_w, img_h = img_size
    layer_w, layer_h = feature_map_size
    if step is not None:
        step_w = step_h = step
    else:
        step_w = img_w / layer_w
        step_h = img_h / layer_h
    level.cells_x, level.cells_y = int(layer_w), int(layer_h)
    # torch.linspace(start, end, n): the Python doubles are cast to fp32 by the kernel (Scalar.to<float>)
    level.x_start, level.x_end = offset[0] * step_w, (offset[0] + layer_w - 1) * step_w
    level.y_start, level.y_end = offset[1] * step_h, (offset[1] + layer_h - 1) * step_h
    n = hws.shape[0]
    if n > N.MAX_BOXES_PER_CELL:
        raise ValueError(f"{n} boxes per cell (at most {N.MAX_BOXES_PER_CELL})")
    level.num_boxes = n
    flat = hws.reshape(-1).tolist()
    for i, v in enumerate(flat):
        level.wh[i] = v
    return level


def _expand_ratios(aspect_ratios, flip: bool) -> List[float]:
    """[r, 1/r for every r > 1] in the order the per-cell boxes are laid out (ssd.py:94-99)."""
    out: List[float] = []
    for r in aspect_ratios:
        if flip and r < 1.0:
            raise AssertionError(f"aspect ratio {r} < 1 cannot be flipped: list ratios >= 1 or pass flip=False")
        out.append(r)
        if flip and r > 1.0:
            out.append(1.0 / r)
    return out


class SsdAnchorGenerator(_AnchorGenerator):
    """Per-level SSD boxes: one box per (flipped) aspect ratio at the level's lower size bound plus, when an upper
    bound is given, the square box at the geometric mean of the two bounds; ``num_branches`` repeats that over a
    linear subdivision of [lower, upper].  Same constructor as detection/anchor_generators/ssd.py:55-110; the bounds
    come either as fractions of the image side (``min_scale`` / ``max_scale``) or in pixels (``min_size`` /
    ``max_size``)."""

    def __init__(self, aspect_ratios, min_scale=None, max_scale=None, min_size=None, max_size=None, step=None,
                 offset=[.5, .5], num_branches=1, flip=True, clip=False):
        super(SsdAnchorGenerator, self).__init__()
        for lower, upper, what in ((min_scale, max_scale, "scale"), (min_size, max_size, "size")):
            if upper is not None and lower is None:
                raise ValueError(f'"max_{what}" should be provided along with "min_{what}"')
        if min_scale is not None and min_size is not None:
            raise ValueError('Either "min_scale" or "min_size" should be provided')
        self.min_scale, self.max_scale, self.min_size, self.max_size = min_scale, max_scale, min_size, max_size
        self.step, self.offset, self.num_branches = step, offset, num_branches
        # ssd.py:147-149 clamps `boxes[..., [0, 2]]` -- an advanced-indexing COPY -- so `clip` never
        # changes the reference's table; kept as an attribute, deliberately without effect
        self.clip = clip
        self.aspect_ratios = _expand_ratios(aspect_ratios, flip)
        self._has_mean_box = bool(max_scale or max_size)
        self.num_ratios = len(self.aspect_ratios) + int(self._has_mean_box)
        self.num_boxes = self.num_ratios * num_branches
        in_pixels = self.min_size is not None and self.max_size is not None
        lo, hi = (self.min_size, self.max_size) if in_pixels else (self.min_scale, self.max_scale)
        bounds = torch.linspace(lo, hi, self.num_branches + 1).unsqueeze(1)          # [branches + 1, 1] fp32
        if in_pixels:
            self.sizes = bounds.expand(-1, 2)
        else:
            self.scales = bounds

    def _shapes(self, img_size) -> torch.Tensor:
        """(w, h) of every box of a cell, fp32 [num_boxes, 2].  The arithmetic keeps the reference's precision steps
        (ssd.py:121-137): an fp32 bound times / over the fp32-rounded double sqrt(r); the mean box is the double
        square root of the fp32 product of the two bounds."""
        if self.min_size is not None and self.max_size is not None:
            bounds = self.sizes
        else:
            bounds = torch.cat([self.scales * img_size[0], self.scales * img_size[1]], dim=1)   # [branches + 1, (w, h)]
        lower, upper = bounds[:-1], bounds[1:]                                                  # [branches, 2] each
        root = torch.tensor([math.sqrt(r) for r in self.aspect_ratios], dtype=torch.float64).to(torch.float32)
        ratio_boxes = torch.stack([lower[:, 0:1] * root, lower[:, 1:2] / root], dim=-1)         # [branches, ratios, 2]
        mean_box = torch.sqrt((lower * upper).to(torch.float64)).to(torch.float32).unsqueeze(1)  # [branches, 1, 2]
        # (an upper bound always exists: without one the bounds' linspace in __init__ has no end point, there as here)
        per_branch = torch.cat([ratio_boxes, mean_box], dim=1)
        return per_branch.reshape(self.num_boxes, 2).contiguous()

    def _level(self, img_size, feature_map_size) -> N.AnchorLevel:
        return _fill_level(N.AnchorLevel(), img_size, feature_map_size, self.step, self.offset, self._shapes(img_size))


class RetinaAnchorGenerator(_AnchorGenerator):
    """detection/anchor_generators/retina_net.py:18-54."""

    def __init__(self, aspect_ratios, level, scale, scales_per_level=1):
        self.aspect_ratios = aspect_ratios
        self.num_boxes = len(aspect_ratios) * scales_per_level
        self.sizes = [scale * (2 ** (level + x / scales_per_level)) for x in range(scales_per_level)]

    def _shapes(self, img_size) -> torch.Tensor:
        hws = torch.empty((self.num_boxes, 2), dtype=torch.float32)
        for j, size in enumerate(self.sizes):
            for i, ar in enumerate(self.aspect_ratios):
                hws[j * len(self.aspect_ratios) + i][0] = size * math.sqrt(ar)
                hws[j * len(self.aspect_ratios) + i][1] = size / math.sqrt(ar)
        return hws

    def _level(self, img_size, feature_map_size) -> N.AnchorLevel:
        return _fill_level(N.AnchorLevel(), img_size, feature_map_size, None, (0.5, 0.5), self._shapes(img_size))


def build_ssd_anchor_generators(num_scales=6, sizes=None, min_scale=None, max_scale=None,
                                aspect_ratios=[[1.0, 2.0]] + [[1.0, 2.0, 3.0]] * 3 + [[1.0, 2.0]] * 2,
                                steps=None, offsets=[0.5, 0.5], num_branches=None, **_ignored):
    """One :class:`SsdAnchorGenerator` per feature level, the levels' bounds taken from a linear scale ramp
    (``min_scale`` .. ``max_scale`` over ``num_scales`` + 1 points) or from explicit pixel ``sizes``
    (ssd.build_anchor_generators, ssd.py:11-53; unknown keys are dropped as ``filter_kwargs`` does)."""
    ramp = min_scale is not None and max_scale is not None
    assert ramp or sizes is not None, "either min_scale + max_scale or sizes"
    per_level = {"aspect_ratios": aspect_ratios, "steps": steps if steps is not None else [None] * num_scales,
                 "num_branches": num_branches if num_branches is not None else [1] * num_scales}
    for name, values in per_level.items():
        assert len(values) == num_scales, f"{name}: {len(values)} entries for {num_scales} levels"
    if ramp:
        edges = torch.linspace(min_scale, max_scale, num_scales + 1)
        logging.info(f'Detector (Scales: {edges[:-1]})')
        bound_keys = ("min_scale", "max_scale")
    else:
        edges, bound_keys = sizes, ("min_size", "max_size")
    return [SsdAnchorGenerator(per_level["aspect_ratios"][i], step=per_level["steps"][i],
                               num_branches=per_level["num_branches"][i],
                               **{bound_keys[0]: edges[i], bound_keys[1]: edges[i + 1]})
            for i in range(num_scales)]


def build_retina_anchor_generators(aspect_ratios, min_level, max_level, scale, scales_per_level, **_ignored):
    """retina_net.build_anchor_generators (retina_net.py:10-16)."""
    return [RetinaAnchorGenerator(aspect_ratios, level, scale, scales_per_level)
            for level in range(min_level, max_level + 1)]


_tables: Dict[Tuple, torch.Tensor] = {}


def generate_anchors(generators: Sequence[_AnchorGenerator], img_size: Tuple[int, int],
                     feature_map_sizes: Sequence[Tuple[int, int]], device: Optional[torch.device] = None,
                     cache: bool = True) -> torch.Tensor:
    """``Detector.generate_anchors`` for all levels at once: ``[A, 4]`` fp32 (cx, cy, w, h) on the GPU.

    ``img_size`` and the feature-map sizes are (width, height), as ``_AnchorGenerator.generate`` passes them."""
    N.require_device()
    if len(generators) != len(feature_map_sizes):
        raise ValueError("one feature-map size per anchor generator")
    if len(generators) > N.MAX_ANCHOR_LEVELS:
        raise ValueError(f"{len(generators)} levels (at most {N.MAX_ANCHOR_LEVELS})")
    device = device or torch.device("cuda", torch.cuda.current_device())
    key = (tuple(id(g) for g in generators), tuple(img_size), tuple(tuple(s) for s in feature_map_sizes), str(device))
    if cache and key in _tables:
        return _tables[key]
    levels = (N.AnchorLevel * max(len(generators), 1))()
    total = 0
    for i, (gen, fm) in enumerate(zip(generators, feature_map_sizes)):
        levels[i] = gen._level(tuple(img_size), tuple(fm))
        total += int(fm[0]) * int(fm[1]) * gen.num_boxes
    out = torch.empty((total, 4), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().ssd_generate_anchors(levels, len(generators), out.data_ptr() if total else None, total,
                                             torch.cuda.current_stream().cuda_stream))
    if cache:
        if len(_tables) >= 16:
            _tables.pop(next(iter(_tables)))
        _tables[key] = out
    return out
