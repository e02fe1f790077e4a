"""Golden values of the reference's mean_average_precision (detection/metrics/mean_average_precision.py)
on seeded synthetic detections.  Build container only: ``python tests/golden/make_golden_map.py``.
Writes ``map.npz`` (inputs + the mAP the REFERENCE returns); pins oracle/map_oracle.py."""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SSD_REFERENCE_ROOT", "/root/reference")
sys.modules.setdefault("jpeg4py", types.SimpleNamespace(JPEG=None))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from detection.metrics.mean_average_precision import mean_average_precision as ref_map  # noqa: E402


def synth_case(gen: torch.Generator, images: int, classes: int, max_gt: int, with_difficult: bool, img: float = 300.0):
    """Ground truth rows (x1,y1,x2,y2,class,score[,difficult]) and detections that hit, nearly hit,
    duplicate and miss them; unique scores (the reference's argsort is unstable on ties)."""
    gts, preds = [], []
    for i in range(images):
        g = int(torch.randint(0, max_gt + 1, (1,), generator=gen))
        centre = torch.rand((g, 2), generator=gen) * img
        side = torch.rand((g, 2), generator=gen) * 0.4 * img + 0.05 * img
        box = torch.cat([centre - side / 2, centre + side / 2], dim=1).clamp_(0, img - 1)
        cls = torch.randint(1, classes + 1, (g, 1), generator=gen).float()
        cols = [box, cls, torch.ones((g, 1))]
        if with_difficult:
            cols.append((torch.rand((g, 1), generator=gen) < 0.25).float())
        gts.append(torch.cat(cols, dim=1).float())
        for r in range(g):
            for _ in range(int(torch.randint(0, 4, (1,), generator=gen))):
                jitter = (torch.rand(4, generator=gen) - 0.5) * side[r].repeat(2) * float(torch.rand(1, generator=gen)) * 1.2
                c = cls[r, 0] if float(torch.rand(1, generator=gen)) < 0.85 else float(torch.randint(1, classes + 1, (1,), generator=gen))
                preds.append(torch.cat([torch.tensor([float(i)]), box[r] + jitter, torch.tensor([float(c)]), torch.rand(1, generator=gen)]))
        for _ in range(int(torch.randint(0, 5, (1,), generator=gen))):
            c2 = torch.rand(2, generator=gen) * img
            s2 = torch.rand(2, generator=gen) * 0.3 * img + 5
            preds.append(torch.cat([torch.tensor([float(i)]), c2 - s2 / 2, c2 + s2 / 2,
                                    torch.randint(1, classes + 2, (1,), generator=gen).float(), torch.rand(1, generator=gen)]))
    preds = torch.stack(preds).float()
    preds[:, 6] = torch.rand(preds.shape[0], generator=gen)          # unique with probability 1
    if not gts[0].shape[0]:
        gts[0] = gts[1].clone() if gts[1].shape[0] else gts[0]
    return preds, gts


def main():
    gen = torch.Generator().manual_seed(23)
    blob, n = {}, 0
    for images, classes, max_gt, diff in [(6, 3, 4, False), (12, 5, 6, True), (40, 20, 8, True), (25, 8, 5, False),
                                          (8, 2, 10, True)]:
        preds, gts = synth_case(gen, images, classes, max_gt, diff)
        labels = {c: str(c) for c in range(0, classes + 3)}
        for voc in (False, True):
            for thr in (0.5, 0.75):
                value = ref_map(preds.clone(), [g.clone() for g in gts], labels, thr, voc=voc, verbose=False)
                cols = gts[0].shape[1]
                flat = torch.cat([g.reshape(-1, cols) for g in gts], dim=0)
                off = np.cumsum([0] + [g.shape[0] for g in gts]).astype(np.int64)
                blob[f"preds_{n}"] = preds.numpy()
                blob[f"gt_flat_{n}"] = flat.numpy()
                blob[f"gt_off_{n}"] = off
                blob[f"cfg_{n}"] = np.array([thr, float(voc)], dtype=np.float64)
                blob[f"map_{n}"] = np.array(value, dtype=np.float64)
                print(n, images, classes, diff, voc, thr, preds.shape[0], value)
                n += 1
    blob["num_cases"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "map.npz"), **blob)


if __name__ == "__main__":
    main()
