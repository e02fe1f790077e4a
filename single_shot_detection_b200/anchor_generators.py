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


class SsdAnchorGenerator(_AnchorGenerator):
    """detection/anchor_generators/ssd.py:55-151."""

    def __init__(self, aspect_ratios, min_scale=None, max_scale=None, min_size=None, max_size=None, step=None,
                 offset=[.5, .5], num_branches=1, flip=True, clip=False):
        super(SsdAnchorGenerator, self).__init__()
        if max_scale is not None and min_scale is None:
            raise ValueError('"max_scale" should be provided along with "min_scale"')
        if max_size is not None and min_size is None:
            raise ValueError('"max_size" should be provided along with "min_size"')
        if min_scale is not None and min_size is not None:
            raise ValueError('Either "min_scale" or "min_size" should be provided')
        self.min_scale, self.max_scale = min_scale, max_scale
        self.min_size, self.max_size = min_size, max_size
        self.num_branches = num_branches
        # ssd.py:147-149 clamps `boxes[..., [0, 2]]` -- an advanced-indexing COPY -- so `clip` never
        # changes the reference's table; kept as an attribute, deliberately without effect
        self.clip = clip
        self.offset = offset
        self.step = step
        self.aspect_ratios = []
        for ar in aspect_ratios:
            assert ar >= 1.0 or not flip
            self.aspect_ratios.append(ar)
            if ar > 1.0 and flip:
                self.aspect_ratios.append(1.0 / ar)
        self.num_ratios = len(self.aspect_ratios)
        if max_scale or max_size:
            self.num_ratios += 1
        self.num_boxes = self.num_ratios * num_branches
        if self.min_size is not None and self.max_size is not None:
            self.sizes = torch.linspace(self.min_size, self.max_size, self.num_branches + 1).unsqueeze(1).expand(-1, 2)
        else:
            self.scales = torch.linspace(self.min_scale, self.max_scale, self.num_branches + 1).unsqueeze(1)

    def _shapes(self, img_size) -> torch.Tensor:
        """(w, h) per box, fp32 [num_boxes, 2] (ssd.py:121-137, same host arithmetic)."""
        img_w, img_h = img_size
        hws = torch.empty((self.num_boxes, 2), dtype=torch.float32)
        if self.min_size is not None and self.max_size is not None:
            sizes = self.sizes
        else:
            sizes = torch.cat([self.scales * img_w, self.scales * img_h], dim=1)
        for j in range(self.num_branches):
            lo, hi = sizes[j], sizes[j + 1]
            i = -1
            for i, r in enumerate(self.aspect_ratios):
                hws[j * self.num_ratios + i][0] = lo[0] * math.sqrt(r)
                hws[j * self.num_ratios + i][1] = lo[1] / math.sqrt(r)
            hws[j * self.num_ratios + i + 1][0] = math.sqrt(lo[0] * hi[0])
            hws[j * self.num_ratios + i + 1][1] = math.sqrt(lo[1] * hi[1])
        return hws

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
    """ssd.build_anchor_generators (ssd.py:11-53; unknown keys are dropped as ``filter_kwargs`` does)."""
    assert sizes is not None or (min_scale is not None and max_scale is not None)
    if steps is None:
        steps = [None] * num_scales
    else:
        assert len(steps) == num_scales
    if num_branches is None:
        num_branches = [1] * num_scales
    else:
        assert len(num_branches) == num_scales
    if min_scale is not None and max_scale is not None:
        scales = torch.linspace(min_scale, max_scale, num_scales + 1)
        logging.info(f'Detector (Scales: {scales[:-1]})')
    else:
        scales = None
    assert len(aspect_ratios) == num_scales
    generators = []
    for i, (ratios, step, branches) in enumerate(zip(aspect_ratios, steps, num_branches)):
        if scales is not None:
            kwargs = {'min_scale': scales[i], 'max_scale': scales[i + 1]}
        else:
            kwargs = {'min_size': sizes[i], 'max_size': sizes[i + 1]}
        generators.append(SsdAnchorGenerator(ratios, step=step, num_branches=branches, **kwargs))
    return generators


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
