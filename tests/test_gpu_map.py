"""GPU parity of the detection metric: csrc/metrics.cu through the C ABI against
the values the reference returned (tests/golden/map.npz) and against oracle/map_oracle.py.

Bars: TP / FP flags bit exact; AP / mAP within 1e-6 relative (fp32 sums; the reference's dot product
is a BLAS call whose summation order is not specified)."""
import math

import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import map_oracle as mo
from test_map_golden import load_case

pytestmark = pytest.mark.gpu

REL = 1e-6


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def _flags_by_index(order, flags, n):
    out = np.full(n, -1, dtype=np.int64)
    out[np.asarray(order, dtype=np.int64)] = np.asarray(flags, dtype=np.int64)
    return out


def _check(preds, gts, thr, voc, dev, want=None):
    from single_shot_detection_b200 import mean_average_precision as M
    res = M.evaluate(preds.to(dev), gts, thr, voc)
    ref, ref_aps, ref_flags = mo.mean_average_precision(preds, gts, thr, voc)
    ref_order = torch.sort(preds[:, 6], descending=True, stable=True).indices.numpy()
    n = preds.shape[0]
    got_flags = _flags_by_index(res.order.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, res.flags.cpu().numpy(), n)
    np.testing.assert_array_equal(got_flags, _flags_by_index(ref_order, ref_flags, n))
    assert sorted(res.per_class) == sorted(ref_aps)
    for c, v in ref_aps.items():
        if math.isnan(v):
            assert math.isnan(res.per_class[c]), c
        else:
            np.testing.assert_allclose(res.per_class[c], v, rtol=REL, atol=1e-7, err_msg=f"class {c}")
    for target in (ref, want):
        if target is None:
            continue
        if math.isnan(target):
            assert math.isnan(res.value)
        else:
            np.testing.assert_allclose(res.value, target, rtol=REL)
    return res


def test_map_matches_reference_golden_values(dev):
    z = gio.load("map.npz")
    for n in range(int(z["num_cases"])):
        preds, gts, thr, voc, want = load_case(z, n)
        _check(preds, gts, thr, voc, dev, want)


def test_reference_signature_accepts_cpu_predictions_like_eval_py(dev):
    from single_shot_detection_b200 import mean_average_precision as M
    z = gio.load("map.npz")
    preds, gts, thr, voc, want = load_case(z, 0)
    labels = {c: str(c) for c in range(32)}
    got = M.mean_average_precision(preds, gts, labels, thr, voc=voc, verbose=False)
    np.testing.assert_allclose(got, want, rtol=REL)


def _synthetic(gen, images, classes, max_gt, difficult, dets_per_image):
    gts, rows = [], []
    for i in range(images):
        g = int(torch.randint(0, max_gt + 1, (1,), generator=gen))
        c = torch.rand((g, 2), generator=gen) * 300
        s = torch.rand((g, 2), generator=gen) * 120 + 15
        box = torch.cat([c - s / 2, c + s / 2], 1).clamp_(0, 299)
        cls = torch.randint(1, classes + 1, (g, 1), generator=gen).float()
        cols = [box, cls, torch.ones((g, 1))]
        if difficult:
            cols.append((torch.rand((g, 1), generator=gen) < 0.2).float())
        gts.append(torch.cat(cols, 1).float())
        for _ in range(dets_per_image):
            if g and float(torch.rand(1, generator=gen)) < 0.7:
                r = int(torch.randint(0, g, (1,), generator=gen))
                b = box[r] + (torch.rand(4, generator=gen) - 0.5) * s[r].repeat(2) * 0.5
                k = cls[r, 0] if float(torch.rand(1, generator=gen)) < 0.9 else float(torch.randint(1, classes + 2, (1,), generator=gen))
            else:
                c2 = torch.rand(2, generator=gen) * 300
                s2 = torch.rand(2, generator=gen) * 100 + 5
                b = torch.cat([c2 - s2 / 2, c2 + s2 / 2])
                k = float(torch.randint(1, classes + 2, (1,), generator=gen))
            rows.append(torch.cat([torch.tensor([float(i)]), b, torch.tensor([float(k)]), torch.rand(1, generator=gen)]))
    return torch.stack(rows).float(), gts


@pytest.mark.parametrize("voc", [False, True])
def test_map_random_vs_oracle(dev, voc):
    gen = torch.Generator().manual_seed(101 + voc)
    for images, classes, max_gt, difficult, per in [(30, 6, 6, True, 20), (64, 20, 10, False, 40), (3, 2, 2, True, 700)]:
        preds, gts = _synthetic(gen, images, classes, max_gt, difficult, per)
        _check(preds, gts, 0.5, voc, dev)


def test_map_tied_scores_keep_input_order(dev):
    """Equal scores: the reference's argsort is unspecified there; oracle and kernel are both stable."""
    gen = torch.Generator().manual_seed(5)
    preds, gts = _synthetic(gen, 12, 3, 5, False, 30)
    preds[:, 6] = (preds[:, 6] * 8).floor() / 8
    _check(preds, gts, 0.5, False, dev)


def test_map_edge_cases(dev):
    from single_shot_detection_b200 import mean_average_precision as M
    gt = [torch.tensor([[10., 10, 50, 50, 1, 1], [100, 100, 150, 160, 2, 1]]), torch.zeros((0, 6))]
    # no detections at all: every class with ground truth scores 0
    res = M.evaluate(torch.zeros((0, 7), device=dev), gt, 0.5)
    assert res.value == 0.0 and res.per_class == {1: 0.0, 2: 0.0}
    # a perfect detection and one of a class that has no ground truth (ignored by the mean)
    preds = torch.tensor([[0., 10, 10, 50, 50, 1, .9], [1, 0, 0, 5, 5, 7, .8], [0, 100, 100, 150, 160, 2, .7],
                          [0, 100, 100, 150, 160, 2, .6]])
    res = _check(preds, gt, 0.5, False, dev)
    assert res.value == 1.0
    # IoU exactly at the threshold is a false positive (`value > iou_threshold`, :62)
    preds = torch.tensor([[0., 10, 10, 50, 30, 1, .9]])
    res = _check(preds, gt, 0.5, False, dev)
    assert res.per_class[1] == 0.0


def test_accumulator_equals_concatenated_predictions(dev):
    """bf/eval.py:54-64: padded batches appended on the device == torch.cat of [index, prediction] rows."""
    from single_shot_detection_b200 import mean_average_precision as M
    gen = torch.Generator().manual_seed(9)
    preds, gts = _synthetic(gen, 24, 5, 6, True, 25)
    T = 32
    acc = M.DetectionAccumulator(capacity=64, device=dev)          # forces growth
    batch = 8
    for b0 in range(0, 24, batch):
        dets = torch.zeros((batch, T, 6))
        counts = torch.zeros((batch,), dtype=torch.int32)
        for i in range(batch):
            rows = preds[preds[:, 0] == b0 + i][:, 1:]
            dets[i, : rows.shape[0]] = rows
            counts[i] = rows.shape[0]
        acc.add(dets.to(dev), counts.to(dev), gts[b0: b0 + batch])
    got = acc.predictions().cpu()
    assert torch.equal(got, preds)                                   # _synthetic emits image-major rows
    value = acc.compute(0.5, voc=True)
    ref, _, _ = mo.mean_average_precision(preds, gts, 0.5, True)
    np.testing.assert_allclose(value, ref, rtol=REL)


def test_map_large_run_properties(dev):
    """COCO-sized evaluation (5000 images x 100 detections): size-independent properties."""
    from single_shot_detection_b200 import mean_average_precision as M
    gen = torch.Generator().manual_seed(77)
    images, classes = 5000, 80
    sizes = torch.randint(1, 9, (images,), generator=gen)
    gts = []
    for i in range(images):
        g = int(sizes[i])
        c = torch.rand((g, 2), generator=gen) * 500
        s = torch.rand((g, 2), generator=gen) * 150 + 20
        gts.append(torch.cat([c - s / 2, c + s / 2, torch.randint(1, classes + 1, (g, 1), generator=gen).float(),
                              torch.ones((g, 1))], 1))
    flat = torch.cat(gts)
    owner = torch.repeat_interleave(torch.arange(images), sizes)
    # detections = every ground-truth box once, exactly (score high) + as many duplicates (score low)
    exact = torch.cat([owner[:, None].float(), flat[:, :4], flat[:, 4:5], 0.5 + 0.5 * torch.rand((flat.shape[0], 1), generator=gen)], 1)
    dup = exact.clone()
    dup[:, 6] = 0.4 * torch.rand((flat.shape[0],), generator=gen)
    preds = torch.cat([exact, dup])[torch.randperm(2 * flat.shape[0], generator=gen)]
    res = M.evaluate(preds.to(dev), gts, 0.5, False)
    flags = res.flags.cpu().numpy()
    # every box is found exactly once, by its high-scoring detection; every duplicate is a false positive
    assert int((flags == 1).sum()) == flat.shape[0] and int((flags == 2).sum()) == flat.shape[0]
    order = res.order.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    assert (preds[order[flags == 1], 6] >= 0.5).all()
    # all true positives rank before all false positives inside a class: AP == 1 for every class
    assert sorted(res.per_class) == list(range(1, classes + 1))
    np.testing.assert_allclose(list(res.per_class.values()), 1.0, rtol=1e-6)
    np.testing.assert_allclose(res.value, 1.0, rtol=1e-6)
    # sortedness of the order the kernel used: class-major, descending score
    cls = preds[order, 5].numpy()
    sc = preds[order, 6].numpy()
    assert (np.diff(cls) >= 0).all()
    assert ((np.diff(sc) <= 0) | (np.diff(cls) > 0)).all()
