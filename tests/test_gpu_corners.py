"""GPU parity of the optional API corners: soft-NMS (stand-alone and inside the
post-processor), generalized IoU, the numpy route of box_utils, match_bipartite -- against the values
the reference produced (tests/golden/corners.npz) and against the oracle.

Bars: picked indices / keep lists exact; scores and boxes 1e-5 relative; GIoU bit exact (add / sub /
mul / div only, every op separately rounded)."""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import anchor_pipeline_oracle as ora
from single_shot_detection_b200 import workloads as wl

pytestmark = pytest.mark.gpu
REL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def test_soft_nms_matches_reference_picks(dev):
    from single_shot_detection_b200 import box_utils
    z = gio.load("corners.npz")
    for i in range(int(z["num_soft"])):
        boxes, scores = torch.from_numpy(z[f"soft_boxes_{i}"]), torch.from_numpy(z[f"soft_scores_{i}"])
        thr, sigma, k = (float(x) for x in z[f"soft_cfg_{i}"])
        (bk, sk), picked = box_utils.nms(boxes.to(dev), scores.to(dev), overlap_threshold=.45, score_threshold=thr,
                                         max_per_class=None if k < 0 else int(k), soft=True, sigma=sigma)
        assert np.array_equal(picked.cpu().numpy(), z[f"soft_picked_{i}"]), i
        want = z[f"soft_out_{i}"]
        assert np.array_equal(torch.cat([bk.reshape(-1, 4), sk.reshape(-1, 1)], 1).cpu().numpy(), want), i


def test_soft_nms_random_vs_oracle(dev):
    from single_shot_detection_b200 import box_utils
    gen = torch.Generator().manual_seed(3)
    for n, thr, sigma in [(3, .1, .5), (33, .05, .5), (200, .3, .25), (400, .5, .5), (512, .45, 2.0)]:
        c = torch.rand((n, 2), generator=gen) * 200
        s = torch.rand((n, 2), generator=gen) * 70 + 10
        boxes = torch.cat([c - s / 2, c + s / 2], 1)
        scores = torch.rand((n,), generator=gen)
        want = ora.gaussian_soft_nms(boxes, scores, thr, sigma)
        (_, sk), picked = box_utils.nms(boxes.to(dev), scores.to(dev), .45, thr, soft=True, sigma=sigma)
        assert torch.equal(picked.cpu(), want), n
        assert torch.equal(sk.cpu(), scores[want])


def test_postprocessor_with_soft_nms_matches_reference(dev):
    from single_shot_detection_b200 import box_coder, postprocessor
    z = gio.load("corners.npz")
    for m in range(int(z["num_post"])):
        w = wl.WORKLOADS[str(z[f"post_workload_{m}"])]
        anchors = wl.build_anchors(w)
        scores, locs = torch.from_numpy(z[f"post_scores_{m}"]), torch.from_numpy(z[f"post_locs_{m}"])
        want = gio.split_ragged(z[f"post_det_flat_{m}"], z[f"post_det_off_{m}"])
        post = postprocessor.Postprocessor(box_coder.BoxCoder(w.xy_scale, w.wh_scale), score_threshold=.2,
                                           nms={"max_per_class": 100, "overlap_threshold": .45, "soft": True, "sigma": .5},
                                           score_converter=w.converter, max_total=40)
        got = post.postprocess((scores.to(dev), locs.to(dev)), anchors)
        assert len(got) == len(want)
        for g, r in zip(got, want):
            assert g.shape == r.shape
            assert torch.equal(g[:, 4].cpu(), r[:, 4])                                   # classes, row by row
            np.testing.assert_allclose(g.cpu().numpy(), r.numpy(), rtol=REL, atol=1e-5)


def test_generalized_iou_bit_exact(dev):
    from single_shot_detection_b200 import box_utils
    z = gio.load("corners.npz")
    a, b = torch.from_numpy(z["giou_a"]).to(dev), torch.from_numpy(z["giou_b"]).to(dev)
    assert np.array_equal(box_utils.generalized_iou(a, b).cpu().numpy(), z["giou_cartesian"])
    assert np.array_equal(box_utils.generalized_iou(a, b[:37], cartesian=False).cpu().numpy(), z["giou_elementwise"])


def test_numpy_route_of_box_utils(dev):
    """bf/preprocessing/functional/box.py:68-69 calls intersection / iou with numpy arrays."""
    from single_shot_detection_b200 import box_utils
    z = gio.load("corners.npz")
    a, b = z["giou_a"], z["giou_b"]
    got = box_utils.iou(a, b)
    assert isinstance(got, np.ndarray) and got.dtype == np.float32
    assert np.array_equal(got, ora.pairwise_iou(torch.from_numpy(a), torch.from_numpy(b)).numpy(), equal_nan=True)
    region = np.array([40., 40, 220, 260], dtype=np.float32)
    inter = box_utils.intersection(region[np.newaxis], a, zero_incorrect=True).squeeze()
    lo = np.maximum(region[:2], a[:, :2])
    hi = np.minimum(region[2:], a[:, 2:])
    want = np.concatenate([lo, hi], axis=1)
    want[(hi < lo).any(axis=1)] = 0
    assert np.array_equal(inter, want)
    elementwise = box_utils.iou(a, inter, cartesian=False)
    assert elementwise.shape == (a.shape[0],)
    (bk, sk), keep = box_utils.nms(a, np.linspace(0.2, 0.9, a.shape[0]).astype(np.float32), .45, .01, max_per_class=20)
    assert isinstance(keep, np.ndarray) and bk.shape[0] == keep.shape[0] == sk.shape[0]
    sc = np.linspace(0.2, 0.9, a.shape[0]).astype(np.float32)
    subset = np.arange(a.shape[0] - 20, a.shape[0])                              # the 20 best (scores increase)
    assert np.array_equal(keep, subset[ora.greedy_nms(a[subset], sc[subset], .45)])


def _giou_module(w):
    import functools
    from single_shot_detection_b200 import box_coder, multibox_loss, sampler
    coder = box_coder.BoxCoder(w.xy_scale, w.wh_scale, w.eps)
    if w.converter == "SOFTMAX":
        smp = functools.partial(sampler.hard_negative_mining, negative_per_positive_ratio=w.ratio,
                                min_negative_per_image=w.min_neg)
        return multibox_loss.MultiboxLoss(smp, coder, {"name": "CrossEntropyLoss"}, {"name": "GeneralizedIoULoss"},
                                          localization_weight=2.0)
    return multibox_loss.MultiboxLoss(sampler.naive_sampler, coder, {"name": "SigmoidFocalLoss", "gamma": 2.0, "alpha": 0.25},
                                      {"name": "GeneralizedIoULoss"}, localization_weight=2.0)


def test_multibox_loss_with_giou_matches_reference_values_and_gradients(dev):
    """MultiboxLoss(localization_loss=GeneralizedIoULoss): (loss, class_loss, loc_loss) and d loss / d locs,
    d loss / d scores against the reference + autograd (tests/golden/corners.npz), 1e-5 relative."""
    z = gio.load("corners.npz")
    for g in range(int(z["num_giou_loss"])):
        w = wl.WORKLOADS[str(z[f"giou_loss_workload_{g}"])]
        anchors = wl.build_anchors(w)
        scores = torch.from_numpy(z[f"giou_loss_scores_{g}"]).to(dev).requires_grad_(True)
        locs = torch.from_numpy(z[f"giou_loss_locs_{g}"]).to(dev).requires_grad_(True)
        target = torch.from_numpy(z[f"giou_loss_target_{g}"]).to(dev)
        before = target.clone()
        loss, class_loss, loc_loss = _giou_module(w)((scores, locs), anchors, target)
        loss.backward()
        assert torch.equal(target, before)                    # the IOU_LOSS branch leaves the target alone (:77-79)
        np.testing.assert_allclose([float(x.detach()) for x in (loss, class_loss, loc_loss)], z[f"giou_loss_values_{g}"], rtol=REL)
        np.testing.assert_allclose(locs.grad.cpu().numpy(), z[f"giou_loss_grad_locs_{g}"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(scores.grad.cpu().numpy(), z[f"giou_loss_grad_scores_{g}"], rtol=1e-4, atol=1e-7)


def test_giou_loss_gradient_random_vs_oracle_autograd(dev):
    from single_shot_detection_b200 import _devcache  # noqa: F401
    from single_shot_detection_b200.ops import OPS
    from single_shot_detection_b200 import _native as N
    gen = torch.Generator().manual_seed(17)
    for name, batch in [("ssd300_voc_b8", 4), ("ssd_mb2_coco_b64", 5)]:
        w = wl.WORKLOADS[name]
        anchors, gt, scores, locs = wl.make_inputs(w, seed=61, batch=batch)
        locs = (locs * 4.0)
        target = ora.assign_targets(gt, anchors, w.matched_threshold, w.unmatched_threshold)
        ref_locs = locs.clone().requires_grad_(True)
        ref = ora.giou_localization_loss(ref_locs, anchors, target, w.xy_scale, w.wh_scale, loc_weight=1.5)
        ref.backward()
        mask = torch.zeros(target.shape[:2], dtype=torch.bool)
        loss3, _, grad_locs = OPS.multibox_loss(scores.to(dev), locs.to(dev), target.to(dev), mask.to(dev),
                                                N.LOSS_SOFTMAX_CE, 0.0, 0.0, 1.0, 1.5, True, anchors.to(dev),
                                                float(w.xy_scale), float(w.wh_scale))
        np.testing.assert_allclose(float(loss3[2]), float(ref), rtol=REL)
        np.testing.assert_allclose(grad_locs.cpu().numpy().reshape(batch, -1), ref_locs.grad.numpy().reshape(batch, -1),
                                   rtol=1e-4, atol=1e-7)


def test_match_bipartite_vs_oracle(dev):
    """detection/matcher.py:7-31 (dead code in the reference, kept importable): greedy one-to-one matching."""
    from single_shot_detection_b200 import matcher
    gen = torch.Generator().manual_seed(8)
    for g, a in [(1, 7), (5, 40), (12, 300)]:
        w = torch.rand((g, a), generator=gen)
        w[torch.rand((g, a), generator=gen) < 0.3] = 0.0
        w[:, 0] += 0.01                                        # every box keeps a positive weight
        ref_box, ref_anchor = ora.greedy_bipartite_match(w.clone())
        box, anchor = matcher.match_bipartite(w.to(dev))
        assert torch.equal(box.cpu(), ref_box) and torch.equal(anchor.cpu(), ref_anchor), (g, a)
    with pytest.raises(AssertionError):
        matcher.match_bipartite(torch.zeros((2, 5), device=dev))


# ----------------------------------------------------------------------------------------------
# box_utils.nms / Postprocessor without a bound on the boxes per class (max_per_class=None or > 512):
# the long-list kernels (csrc/nms_large.cu) against torchvision / the oracle
# ----------------------------------------------------------------------------------------------
def _random_boxes(n, gen, span=400.0):
    c = torch.rand((n, 2), generator=gen) * span
    s = torch.rand((n, 2), generator=gen) * 60 + 1
    return torch.cat([c - s / 2, c + s / 2], dim=1)


@pytest.mark.parametrize("n,k,thr", [(3000, None, 0.45), (513, None, 0.5), (5000, 600, 0.3), (2049, 2049, 0.6), (4100, None, 0.0)])
def test_nms_long_lists_vs_torchvision(dev, n, k, thr):
    import torchvision
    from single_shot_detection_b200 import box_utils
    gen = torch.Generator().manual_seed(n)
    boxes = _random_boxes(n, gen)
    scores = torch.rand((n,), generator=gen)
    scores[n // 2:n // 2 + 200] = scores[:200]                     # exact score ties
    boxes[n // 3:n // 3 + 50] = boxes[:50]                         # duplicate boxes
    (bk, sk), keep = box_utils.nms(boxes.to(dev), scores.to(dev), thr, 0.01, max_per_class=k)
    if k is None or k >= n:
        ref = torchvision.ops.nms(boxes, scores, thr)
    else:                                                          # bf/utils/box_utils.py:186-188 then :193
        sub = ora.select_top_scores(scores, k, canonical=True)
        ref = sub[torchvision.ops.nms(boxes[sub], scores[sub], thr)]
    assert keep.cpu().tolist() == ref.tolist()
    assert torch.equal(bk.cpu(), boxes[ref]) and torch.equal(sk.cpu(), scores[ref])


def test_soft_nms_long_list_vs_oracle(dev):
    from single_shot_detection_b200 import box_utils
    gen = torch.Generator().manual_seed(5)
    n = 700
    boxes = _random_boxes(n, gen, span=250.0)
    scores = torch.rand((n,), generator=gen) * 0.9 + 0.05
    (bk, sk), keep = box_utils.nms(boxes.to(dev), scores.to(dev), 0.45, 0.3, max_per_class=None, soft=True, sigma=0.5)
    ref = ora.gaussian_soft_nms(boxes, scores, 0.3, 0.5)
    assert keep.cpu().tolist() == ref.tolist()
    assert torch.equal(sk.cpu(), scores[ref])


@pytest.mark.parametrize("name,k", [("tiny_voc_b3", None), ("tiny_sigmoid_b2", None), ("ssd300_voc_b8", 600)])
def test_postprocessor_without_class_bound_vs_oracle(dev, name, k):
    """Postprocessor(nms={'max_per_class': None | > 512}): detection/postprocessor.py:57-76 over every box above
    the threshold."""
    from single_shot_detection_b200.box_coder import BoxCoder
    from single_shot_detection_b200.postprocessor import Postprocessor
    w = wl.WORKLOADS[name]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=31, batch=2)
    thr = 0.2 if k is None else w.score_threshold
    post = Postprocessor(BoxCoder(w.xy_scale, w.wh_scale, w.eps), thr,
                         {"max_per_class": k, "overlap_threshold": w.overlap_threshold},
                         score_converter=w.converter, max_total=w.max_total)
    dets = post.postprocess((scores.to(dev), locs.to(dev)), anchors)
    ref = ora.postprocess(scores, locs, anchors, xy_scale=w.xy_scale, wh_scale=w.wh_scale, score_threshold=thr,
                          overlap_threshold=w.overlap_threshold, max_per_class=k, max_total=w.max_total,
                          converter=w.converter, canonical=True, use_torchvision=True)
    assert len(dets) == len(ref)
    for d, r in zip(dets, ref):
        d = d.cpu()
        assert d.shape == r.shape
        assert torch.equal(d[:, 4], r[:, 4])
        torch.testing.assert_close(d[:, 5], r[:, 5], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(d[:, :4], r[:, :4], rtol=1e-5, atol=2e-5 * w.img)


def test_device_copy_cache_never_serves_a_stale_table(dev):
    """The reference regenerates its CPU anchors every step: a new tensor may reuse the freed one's address."""
    from single_shot_detection_b200 import _devcache
    for _ in range(8):
        a = torch.rand((300, 4))
        d = _devcache.device_copy(a, dev)
        assert torch.equal(d.cpu(), a)
        assert _devcache.device_copy(a, dev) is d               # same object, unmodified: cached
        a.mul_(2.0)                                             # modified in place: a new copy
        assert torch.equal(_devcache.device_copy(a, dev).cpu(), a)
        b = a.clone()
        assert _devcache.device_copy(b, dev) is _devcache.device_copy(a, dev)    # equal content: one device copy
        del a, b
