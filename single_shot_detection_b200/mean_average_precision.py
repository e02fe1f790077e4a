"""mean_average_precision -- same interface as the reference's
``detection/metrics/mean_average_precision.py`` -- and the device-side accumulation that replaces the
loop of ``bf/eval.py:54-70``.

The reference concatenates every batch's detections on the device, copies them to the CPU and walks
them in a Python loop (one ``box_utils.iou`` call per detection).  Here the detections never leave
the GPU: :class:`DetectionAccumulator` compacts the post-processor's padded ``[B, T, 6]`` output into
``[N, 7]`` rows batch after batch (``ssd_map_append``), and ``ssd_mean_average_precision`` turns the
walk into a sort, one matching pass with ``atomicMin`` and one scan per class (csrc/metrics.cu).
The only device->host traffic is the result: mAP + the per-class values.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence

import torch

from . import _native as N
from .ops import workspace
from .target_assigner import pack_ground_truth

DIFFICULT_INDEX = 6            # bf/datasets/detection_dataset.py


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class MapResult:
    """Device-side outputs of one evaluation (``flags`` / ``order`` are in class-major,
    descending-score order; see include/ssd_b200.h)."""

    def __init__(self, value: float, per_class: Dict[int, float], flags: torch.Tensor, order: torch.Tensor):
        self.value, self.per_class, self.flags, self.order = value, per_class, flags, order


def evaluate(predictions: torch.Tensor, gts: Sequence[torch.Tensor], iou_threshold: float, voc: bool = False,
             num_classes: Optional[int] = None) -> MapResult:
    """``predictions`` [N, 7] fp32 on the GPU (image id, box, class, score); ``gts`` the per-image
    ground-truth list.  One host sync: the read of the result."""
    N.require_device()
    if not predictions.is_cuda:
        raise TypeError("mean_average_precision needs CUDA predictions (no CPU fallback)")
    device = predictions.device
    preds = predictions if predictions.dtype == torch.float32 else predictions.float()
    preds = preds if preds.is_contiguous() else preds.contiguous()
    count = int(preds.shape[0])
    use_difficult = gts[0].size(1) > DIFFICULT_INDEX                      # mean_average_precision.py:22
    if num_classes is None:                                               # the ground truth is host data in bf/eval.py
        top = [int(g[:, 4].max()) for g in gts if g.shape[0]]
        num_classes = (max(top) + 1) if top else 0
    num_classes = max(int(num_classes), 0)
    with torch.cuda.device(device):
        packed = pack_ground_truth(gts, device)
        total_gt = int(packed.rows.shape[0])
        if use_difficult and total_gt and packed.rows.shape[1] <= DIFFICULT_INDEX:
            raise ValueError("ground truth mixes rows with and without the difficult column")
        gt_cols = int(packed.rows.shape[1]) if total_gt else (7 if use_difficult else 6)
        ap = torch.empty((max(num_classes, 1),), dtype=torch.float32, device=device)
        out = torch.empty((2,), dtype=torch.float64, device=device)
        flags = torch.empty((max(count, 1),), dtype=torch.uint8, device=device)
        order = torch.empty((max(count, 1),), dtype=torch.int32, device=device)
        ws_bytes = N.lib().ssd_map_workspace_bytes(count, total_gt, num_classes)
        ws = workspace(ws_bytes, device, "map")
        N.check(N.lib().ssd_mean_average_precision(
            preds.data_ptr() if count else None, count, packed.rows.data_ptr() if total_gt else None, gt_cols,
            packed.offsets.data_ptr(), len(gts), total_gt, num_classes, float(iou_threshold), int(use_difficult),
            int(bool(voc)), ap.data_ptr(), out.data_ptr(), flags.data_ptr(), order.data_ptr(), ws.data_ptr(),
            ws.numel(), _stream()))
        host_ap = ap[:num_classes].cpu()
        host_out = out.cpu()                                              # syncs
    per_class = {c: float(v) for c, v in enumerate(host_ap.tolist()) if v != -1.0}
    return MapResult(float(host_out[0]), per_class, flags[:count], order[:count])


def mean_average_precision(predictions, gts, class_labels, iou_threshold, voc=False, verbose=True):
    """
    Args:
        predictions: torch.tensor(:shape [NumBoxes, 7] ~ {[0] - image_id, [1-4] - box, [5] - class, [6] - score})
        gts: list(:len NumImages) ~ torch.tensor(:shape [NumBoxes_i, NumAttributes])
        class_labels: dict(:keys ClassId, :values ClassName)
        iou_threshold: float
        voc: bool
        verbose: bool
    Returns:
        mAP: float

    ``bf/eval.py:66`` hands the metric a CPU copy of the detections; a CPU tensor is shipped back with
    one copy (prefer :class:`DetectionAccumulator`, which never brings them to the host).
    """
    if not predictions.is_cuda:
        predictions = predictions.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    result = evaluate(predictions, gts, iou_threshold, voc)
    if not result.per_class:
        raise ZeroDivisionError("division by zero")                      # sum(...) / len({}) in the reference (:115)
    if verbose:
        logging.info('Mean Average Precision results:')
        for class_index in sorted(result.per_class):
            logging.info(f'{class_labels[class_index]}: {result.per_class[class_index]:6f}')
        logging.info(f'Total mean: {result.value:6f}')
    return result.value


class DetectionAccumulator:
    """Keeps an evaluation run's detections on the device (the ``predictions`` / ``ground_truths``
    lists of ``bf/eval.py:40-59``).

    ``add(dets, counts, ground_truth)`` takes the padded output of ``Postprocessor.postprocess_padded``
    -- no sync, one launch; ``compute`` evaluates everything added so far."""

    def __init__(self, capacity: int = 1 << 16, device: Optional[torch.device] = None):
        N.require_device()
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.rows = torch.empty((int(capacity), 7), dtype=torch.float32, device=self.device)
        self._cursor = torch.zeros((2, 2), dtype=torch.int64, device=self.device)     # ping-pong
        self._slot = 0
        self._upper = 0                          # host-side upper bound of the rows used
        self.images = 0
        self.ground_truths: List[torch.Tensor] = []

    def reset(self) -> None:
        self._cursor.zero_()
        self._slot = self._upper = self.images = 0
        self.ground_truths = []

    def _grow(self, need: int) -> None:
        if need <= self.rows.shape[0]:
            return
        bigger = torch.empty((max(need, 2 * self.rows.shape[0]), 7), dtype=torch.float32, device=self.device)
        bigger[: self.rows.shape[0]].copy_(self.rows)
        self.rows = bigger

    def add(self, dets: torch.Tensor, counts: torch.Tensor, ground_truth: Sequence[torch.Tensor]) -> None:
        batch, max_total = int(dets.shape[0]), int(dets.shape[1])
        if len(ground_truth) != batch:
            raise ValueError("one ground-truth tensor per image")
        if dets.dtype != torch.float32 or not dets.is_contiguous() or dets.shape[2] != 6 or counts.dtype != torch.int32:
            raise TypeError("expects the padded [B, T, 6] fp32 detections and int32 counts of postprocess_padded")
        self._grow(self._upper + batch * max_total)           # bound without reading the counts back
        cur_in, cur_out = self._cursor[self._slot], self._cursor[1 - self._slot]
        with torch.cuda.device(self.device):
            N.check(N.lib().ssd_map_append(dets.data_ptr(), counts.data_ptr(), batch, max_total, self.images,
                                           self.rows.data_ptr(), self.rows.shape[0], cur_in.data_ptr(),
                                           cur_out.data_ptr(), _stream()))
        self._slot = 1 - self._slot
        self._upper += batch * max_total
        self.images += batch
        self.ground_truths += list(ground_truth)

    def count(self) -> int:
        """Rows accumulated so far (one host sync)."""
        n = int(self._cursor[self._slot, 0].item())
        self._upper = n
        return n

    def predictions(self) -> torch.Tensor:
        """The ``[N, 7]`` tensor ``bf/eval.py:64`` builds with ``torch.cat`` (a view, on the device)."""
        return self.rows[: self.count()]

    def compute(self, iou_threshold: float, voc: bool = False, class_labels=None, verbose: bool = False) -> float:
        result = evaluate(self.predictions(), self.ground_truths, iou_threshold, voc)
        self.last_result = result
        if verbose and class_labels is not None:
            for class_index in sorted(result.per_class):
                logging.info(f'{class_labels[class_index]}: {result.per_class[class_index]:6f}')
        return result.value
