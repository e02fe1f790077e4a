"""Anchor (prior box) tables for the benchmark / test workloads.

The anchor pipeline takes ``anchors[A, 4]`` = (cx, cy, w, h) in pixels as a constant input.  The
reference builds it on the CPU once per (image size, feature-map sizes) with
``detection/anchor_generators/ssd.py:55-151`` and ``retina_net.py:18-54`` and concatenates the
levels in ``detection/detector.py:82-86``.  The GPU box has no copy of the reference, so the
tables are rebuilt here; ``tests/test_oracle_golden.py`` checks them bit-for-bit against tables
the reference's own generators produced (``tests/golden/anchors_*.npz``).

Host-side, fp32, computed once: this is input preparation, not part of the timed path.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch


def _cell_centres(img_extent: float, cells: int, step, offset: float) -> torch.Tensor:
    if step is None:
        step = img_extent / cells
    return torch.linspace(offset * step, (offset + cells - 1) * step, cells)


def _level_table(centres_x: torch.Tensor, centres_y: torch.Tensor, shapes: torch.Tensor) -> torch.Tensor:
    """[H, W, nb, 4] laid out as the detector heads emit them (row-major cells, shapes fastest)."""
    h, w, nb = centres_y.numel(), centres_x.numel(), shapes.shape[0]
    table = torch.empty((h, w, nb, 4), dtype=torch.float32)
    table[..., 0] = centres_x.view(1, w, 1)
    table[..., 1] = centres_y.view(h, 1, 1)
    table[..., 2] = shapes[:, 0]
    table[..., 3] = shapes[:, 1]
    return table


def ssd_level_shapes(lo: torch.Tensor, hi: torch.Tensor, ratios: Sequence[float]) -> torch.Tensor:
    """(w, h) of every box of one SSD level.  ``lo``/``hi`` are fp32 [2] = (size_w, size_h) of
    this and the next scale.  Ratios >1 are followed by their reciprocal; the last box is the
    geometric mean of the two scales (ssd.py:83-97, 128-137)."""
    expanded: List[float] = []
    for r in ratios:
        expanded.append(r)
        if r > 1.0:
            expanded.append(1.0 / r)
    shapes = torch.empty((len(expanded) + 1, 2), dtype=torch.float32)
    for k, r in enumerate(expanded):
        shapes[k, 0] = lo[0] * math.sqrt(r)
        shapes[k, 1] = lo[1] / math.sqrt(r)
    shapes[len(expanded), 0] = math.sqrt(lo[0] * hi[0])
    shapes[len(expanded), 1] = math.sqrt(lo[1] * hi[1])
    return shapes


def ssd_anchor_table(img_size: Tuple[int, int], fmap_sizes: Sequence[Tuple[int, int]],
                     min_scale: float, max_scale: float,
                     aspect_ratios: Sequence[Sequence[float]],
                     offset: Tuple[float, float] = (0.5, 0.5)) -> torch.Tensor:
    """SSD anchors for every level, [A, 4] fp32 (cx, cy, w, h) in pixels.

    ``img_size`` and every ``fmap_sizes`` entry are (width, height).  Scale schedule:
    ``linspace(min_scale, max_scale, levels + 1)`` (ssd.py:34-36); one branch per level.
    """
    levels = len(fmap_sizes)
    assert len(aspect_ratios) == levels
    img_w, img_h = img_size
    schedule = torch.linspace(min_scale, max_scale, levels + 1)
    parts = []
    for lvl, (fw, fh) in enumerate(fmap_sizes):
        pair = torch.linspace(schedule[lvl], schedule[lvl + 1], 2).unsqueeze(1)
        sizes = torch.cat([pair * img_w, pair * img_h], dim=1)       # [2, 2]
        shapes = ssd_level_shapes(sizes[0], sizes[1], aspect_ratios[lvl])
        xs = _cell_centres(img_w, fw, None, offset[0])
        ys = _cell_centres(img_h, fh, None, offset[1])
        parts.append(_level_table(xs, ys, shapes).view(-1))
    return torch.cat(parts).view(-1, 4)


def retina_anchor_table(img_size: Tuple[int, int], fmap_sizes: Sequence[Tuple[int, int]],
                        aspect_ratios: Sequence[float], min_level: int, scale: float,
                        scales_per_level: int) -> torch.Tensor:
    """RetinaNet anchors (retina_net.py:18-54): level L boxes have side scale*2^(L + k/spl)."""
    img_w, img_h = img_size
    parts = []
    for i, (fw, fh) in enumerate(fmap_sizes):
        level = min_level + i
        sides = [scale * (2 ** (level + k / scales_per_level)) for k in range(scales_per_level)]
        shapes = torch.empty((len(sides) * len(aspect_ratios), 2), dtype=torch.float32)
        for j, side in enumerate(sides):
            for k, ar in enumerate(aspect_ratios):
                shapes[j * len(aspect_ratios) + k, 0] = side * math.sqrt(ar)
                shapes[j * len(aspect_ratios) + k, 1] = side / math.sqrt(ar)
        xs = _cell_centres(img_w, fw, None, 0.5)
        ys = _cell_centres(img_h, fh, None, 0.5)
        parts.append(_level_table(xs, ys, shapes).view(-1))
    return torch.cat(parts).view(-1, 4)
