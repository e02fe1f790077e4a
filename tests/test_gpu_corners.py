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
