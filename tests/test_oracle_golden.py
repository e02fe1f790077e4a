"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz).

Runs without a GPU.  Bit-exact (torch.equal) for everything that is index / copy / plain fp32
arithmetic; the transcendental paths (log_softmax, softmax, exp, log) go through the same torch
CPU kernels in oracle and reference, so they are compared at 1e-6 to allow a different CPU ISA
dispatch on another host.
"""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import anchor_pipeline_oracle as ora
from single_shot_detection_b200 import workloads as wl


@pytest.fixture(scope="module", params=gio.PIPELINE_CASES)
def case(request):
    return gio.PipelineCase(request.param)


def test_anchor_tables_match_reference_generators():
    z = gio.load("anchors.npz")
    assert len(z.files) == 8
    for name in z.files:
        mine = wl.build_anchors(wl.WORKLOADS[name])
        assert torch.equal(mine, torch.from_numpy(z[name])), name


def test_assign_targets_bit_exact(case):
    target, matches = ora.assign_targets(case.gt, case.anchors, case.matched, case.unmatched,
                                         return_match=True)
    assert torch.equal(target, case.target)
    assert torch.equal(torch.stack(matches), case.match_idx)
    assert not ora.positive_rows_have_nan(target)


def test_iou_bit_exact(case):
    if not case.full or case.gt[0].shape[0] == 0:
        pytest.skip("no IoU stored for this case")
    iou = ora.pairwise_iou(case.gt[0][:, :4], ora.corners_from_centroids(case.anchors))
    assert torch.equal(iou, case.t("iou0"))


def test_samplers(case):
    cls = case.target[..., 4].long()
    logits = case.scores.view(case.B, case.A, case.C)
    assert torch.equal(ora.positives_mask(cls), case.naive_mask)
    # the reference's own (unstable-sort) ranking, same torch ops -> same mask
    ref_like = ora.mine_hard_negatives(logits, cls, case.w.ratio, case.w.min_neg, canonical=False)
    assert torch.equal(ref_like, case.hnm_mask)
    # canonical tie rule selects the same set unless a loss value ties across the cut
    loss = ora.background_loss(logits)
    canon = ora.mine_hard_negatives(logits, cls, case.w.ratio, case.w.min_neg, canonical=True)
    tied = ora.mining_boundary_tie(loss, cls, case.w.ratio, case.w.min_neg)
    for i in range(case.B):
        if not tied[i]:
            assert torch.equal(canon[i], case.hnm_mask[i]), i
        else:                       # only the count is defined
            assert int(canon[i].sum()) == int(case.hnm_mask[i].sum())
    if case.full:
        torch.testing.assert_close(loss, case.t("neg_loss"), rtol=1e-6, atol=1e-6)


def test_box_coding(case):
    # in-place route used by the loss (multibox_loss.py:81-82)
    t = case.target.clone()
    tl = t[..., 0:4]
    ora.centroids_from_corners(tl, inplace=True)
    if case.full:
        assert torch.equal(tl, case.t("centroids_inplace"))
    ora.encode_boxes(tl, case.anchors, case.w.xy_scale, case.w.wh_scale, case.w.eps, inplace=True)
    torch.testing.assert_close(case.enc_view(tl), case.enc_inplace, rtol=1e-6, atol=1e-6, equal_nan=True)
    assert torch.equal(case.enc_view(tl)[..., :2], case.enc_inplace[..., :2])          # xy has no transcendental
    if not case.full:
        return
    cen = ora.centroids_from_corners(case.target[..., 0:4])
    assert torch.equal(cen, case.t("centroids_oop"))
    enc = ora.encode_boxes(cen, case.anchors, case.w.xy_scale, case.w.wh_scale, case.w.eps)
    torch.testing.assert_close(enc, case.t("enc_oop"), rtol=1e-6, atol=1e-6, equal_nan=True)
    locs = case.locs.view(case.B, case.A, 4)
    dec = ora.decode_boxes(locs, case.anchors, case.w.xy_scale, case.w.wh_scale)
    torch.testing.assert_close(dec, case.t("decoded"), rtol=1e-6, atol=1e-6)
    assert torch.equal(dec[..., :2], case.t("decoded")[..., :2])
    dec_in = ora.decode_boxes(locs.clone(), case.anchors, case.w.xy_scale, case.w.wh_scale, inplace=True)
    torch.testing.assert_close(dec_in, case.t("decoded_inplace"), rtol=1e-6, atol=1e-6)
    assert torch.equal(ora.corners_from_centroids(case.t("decoded")), case.t("decoded_corners"))


def _same_detections(mine, ref, exact):
    assert len(mine) == len(ref)
    for m, r in zip(mine, ref):
        assert m.shape == r.shape
        if exact:
            assert torch.equal(m, r)
        else:
            torch.testing.assert_close(m, r, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("max_total", ["cfg", None])
def test_postprocess_end_to_end(case, max_total):
    w = case.w
    mt = w.max_total if max_total == "cfg" else None
    ref = case.dets if max_total == "cfg" else case.dets_all
    # reference tie behaviour + torchvision NMS: the very same torch ops
    mine = ora.postprocess(case.scores, case.locs, case.anchors, xy_scale=w.xy_scale,
                           wh_scale=w.wh_scale, score_threshold=w.score_threshold,
                           overlap_threshold=w.overlap_threshold, max_per_class=w.max_per_class,
                           max_total=mt, converter=w.converter, canonical=False, use_torchvision=True)
    _same_detections(mine, ref, exact=False)
    # canonical ties + the restated NMS: same answer when no score ties at a selection boundary
    probs = ora.convert_scores(case.scores.view(case.B, case.A, case.C), w.converter)
    tie = ora.class_topk_boundary_tie(probs, w.score_threshold, w.max_per_class)
    if bool(tie.any()):
        pytest.skip("score tie at a top-k boundary: the reference answer is not unique")
    canon = ora.postprocess(case.scores, case.locs, case.anchors, xy_scale=w.xy_scale,
                            wh_scale=w.wh_scale, score_threshold=w.score_threshold,
                            overlap_threshold=w.overlap_threshold, max_per_class=w.max_per_class,
                            max_total=mt, converter=w.converter, canonical=True, use_torchvision=False)
    for m, r in zip(canon, ref):
        assert m.shape == r.shape
        # the final top-k (sorted=True) may order equal scores differently: compare as sorted rows
        if mt is not None and not torch.equal(m, r):
            ms = m[torch.argsort(m[:, 5], descending=True, stable=True)]
            assert len(torch.unique(r[:, 5])) < r.shape[0] or torch.allclose(ms, r, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(torch.sort(m[:, 5])[0], torch.sort(r[:, 5])[0], rtol=1e-6, atol=1e-6)


def test_keep_anchors_of_config_cases(case):
    """BASELINE-config fixtures: the oracle keeps the same anchors, class by class and in the same order, as the
    reference did (the anchor of every kept row was recovered from the reference's own intermediates)."""
    if case.det_all_anchor is None:
        pytest.skip("no kept-anchor list in this fixture")
    w = case.w
    fg = ora.convert_scores(case.scores.view(case.B, case.A, case.C), w.converter)      # foreground columns
    if bool(ora.class_topk_boundary_tie(fg, w.score_threshold, w.max_per_class).any()):
        pytest.skip("score tie at a top-k boundary: the reference answer is not unique")
    corners = ora.corners_from_centroids(ora.decode_boxes(case.locs.view(case.B, case.A, 4), case.anchors,
                                                          w.xy_scale, w.wh_scale))
    dets, keep = ora.detections_from_scores(fg, corners, w.score_threshold, w.overlap_threshold, w.max_per_class,
                                            None, canonical=True, use_torchvision=False, return_keep=True)
    for i in range(case.B):
        assert torch.cat(keep[i]).tolist() == case.det_all_anchor[i].tolist(), i
        torch.testing.assert_close(dets[i], case.dets_all[i], rtol=1e-6, atol=1e-6)


def test_postprocess_stage_exact(case):
    """Selection + NMS on the reference's own probabilities / decoded corners is bit exact."""
    if not case.full:
        pytest.skip("intermediates not stored for the large cases")
    w = case.w
    probs = case.t("probs")
    fg = probs[..., 1:] if w.converter == "SOFTMAX" else probs
    corners = case.t("decoded_corners")
    tie = ora.class_topk_boundary_tie(fg, w.score_threshold, w.max_per_class)
    assert not bool(tie.any())
    mine = ora.detections_from_scores(fg, corners, w.score_threshold, w.overlap_threshold,
                                      w.max_per_class, None, canonical=True, use_torchvision=False)
    _same_detections(mine, case.dets_all, exact=True)


def test_loss_with_oracle_sampler_and_coder(case):
    w = case.w
    cls = case.target[..., 4].long()
    t = case.target.clone()
    tl = t[..., 0:4]
    ora.centroids_from_corners(tl, inplace=True)
    ora.encode_boxes(tl, case.anchors, w.xy_scale, w.wh_scale, w.eps, inplace=True)
    if w.converter == "SOFTMAX":
        mask = ora.mine_hard_negatives(case.scores.view(case.B, case.A, case.C), cls, w.ratio, w.min_neg,
                                       canonical=False)
        loss3 = ora.multibox_loss_ce_smoothl1(case.scores, case.locs, case.anchors, case.target, mask, tl)
    else:       # the sigmoid configs: naive sampler + SigmoidFocalLoss(gamma=2, alpha=.25), make_golden.py
        loss3 = ora.multibox_loss_focal_smoothl1(case.scores, case.locs, case.target, ora.positives_mask(cls), tl)
    np.testing.assert_allclose([float(x) for x in loss3], case.loss3, rtol=1e-5)


def test_greedy_nms_matches_torchvision_golden():
    z = gio.load("nms.npz")
    for i in range(int(z["num_cases"])):
        keep = ora.greedy_nms(z[f"boxes_{i}"], z[f"scores_{i}"], float(z[f"thr_{i}"]))
        assert np.array_equal(keep, z[f"keep_{i}"]), i
    for i in range(int(z["num_topk_cases"])):
        bx, sc = torch.from_numpy(z[f"topk_boxes_{i}"]), torch.from_numpy(z[f"topk_scores_{i}"])
        (bk, sk), keep, subset = ora.class_nms(bx, sc, 0.45, int(z[f"topk_k_{i}"]))
        assert torch.equal(bk, torch.from_numpy(z[f"topk_kept_boxes_{i}"]))
        assert torch.equal(sk, torch.from_numpy(z[f"topk_kept_scores_{i}"]))


def test_greedy_nms_matches_installed_torchvision_random():
    """torchvision travels with the image, so the restated NMS is also checked live."""
    import torchvision
    gen = torch.Generator().manual_seed(5)
    for n in (0, 1, 2, 33, 100, 150):
        for thr in (0.0, 0.3, 0.45, 0.5, 0.999):
            c = torch.rand((n, 2), generator=gen) * 50
            s = torch.rand((n, 2), generator=gen) * 30 + 0.5
            boxes = torch.cat([c - s / 2, c + s / 2], dim=1)
            scores = (torch.rand((n,), generator=gen) * 16).round() / 16      # many exact ties
            ref = torchvision.ops.nms(boxes, scores, thr).numpy()
            assert np.array_equal(ora.greedy_nms(boxes.numpy(), scores.numpy(), thr), ref), (n, thr)


def test_matcher_tie_rules_known_answers():
    # SURVEY.md §7 hard part 4
    iou = torch.tensor([[0.6, 0.6, 0.0, 0.2],
                        [0.6, 0.7, 0.0, 0.2],
                        [0.0, 0.0, 0.0, 0.0]])
    idx = ora.match_anchors(iou, 0.5, 0.3)
    # forced: gt0 -> anchor0 (lowest index of the 0.6 tie), gt1 -> anchor1, gt2 (all zero) -> anchor0,
    # the collision on anchor 0 is won by the highest GT index
    assert idx.tolist() == [2, 1, -2, -2]
    idx = ora.match_anchors(torch.tensor([[0.4, 0.35, 0.1]]), 0.5, 0.3)
    assert idx.tolist() == [0, -1, -2]
    # fp32 threshold compare: fp32(0.4) < 0.4 is False
    idx = ora.match_anchors(torch.tensor([[0.9, np.float32(0.4)]]), 0.4, 0.4)
    assert idx.tolist() == [0, 0]
    gt, an = ora.greedy_bipartite_match(torch.tensor([[0.9, 0.8, 0.1], [0.85, 0.1, 0.2]]))
    assert an.tolist() == [0, 2]


# ----------------------------------------------------------------------------------------------
# optional API corners: soft-NMS, generalized IoU -- tests/golden/corners.npz
# ----------------------------------------------------------------------------------------------
def test_soft_nms_oracle_matches_reference_picks():
    z = gio.load("corners.npz")
    for i in range(int(z["num_soft"])):
        boxes, scores = torch.from_numpy(z[f"soft_boxes_{i}"]), torch.from_numpy(z[f"soft_scores_{i}"])
        thr, sigma, k = (float(x) for x in z[f"soft_cfg_{i}"])
        assert k < 0 or k >= boxes.shape[0]                       # no unsorted top-k in the fixtures
        picked = ora.gaussian_soft_nms(boxes, scores, thr, sigma)
        assert np.array_equal(picked.numpy(), z[f"soft_picked_{i}"]), i
        out = torch.cat([boxes[picked].reshape(-1, 4), scores[picked].reshape(-1, 1)], dim=1)
        assert np.array_equal(out.numpy(), z[f"soft_out_{i}"]), i
    # the loop head tests the SUM of the remaining indices: a lone box (index 0) is never picked
    assert z["soft_picked_0"].size == 0


def test_soft_postprocess_oracle_matches_reference():
    z = gio.load("corners.npz")
    for m in range(int(z["num_post"])):
        w = wl.WORKLOADS[str(z[f"post_workload_{m}"])]
        anchors = wl.build_anchors(w)
        scores, locs = torch.from_numpy(z[f"post_scores_{m}"]), torch.from_numpy(z[f"post_locs_{m}"])
        want = gio.split_ragged(z[f"post_det_flat_{m}"], z[f"post_det_off_{m}"])
        got = ora.postprocess(scores, locs, anchors, xy_scale=w.xy_scale, wh_scale=w.wh_scale, score_threshold=.2,
                              overlap_threshold=.45, max_per_class=100, max_total=40, converter=w.converter,
                              canonical=False, soft_sigma=.5)
        for g, r in zip(got, want):
            assert g.shape == r.shape
            np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=1e-6, atol=1e-6)


def test_generalized_iou_oracle_matches_reference():
    z = gio.load("corners.npz")
    a, b = torch.from_numpy(z["giou_a"]), torch.from_numpy(z["giou_b"])
    assert np.array_equal(ora.generalized_iou(a, b).numpy(), z["giou_cartesian"])
    assert np.array_equal(ora.generalized_iou(a, b[:37], cartesian=False).numpy(), z["giou_elementwise"])
    assert z["giou_cartesian"][3, 5] == 1.0


def test_giou_localization_loss_oracle_matches_reference_autograd():
    """Value and gradient of the GIoU localisation term against the reference's MultiboxLoss + autograd."""
    z = gio.load("corners.npz")
    for g in range(int(z["num_giou_loss"])):
        w = wl.WORKLOADS[str(z[f"giou_loss_workload_{g}"])]
        anchors = wl.build_anchors(w)
        locs = torch.from_numpy(z[f"giou_loss_locs_{g}"]).clone().requires_grad_(True)
        target = torch.from_numpy(z[f"giou_loss_target_{g}"])
        loc_loss = ora.giou_localization_loss(locs, anchors, target, w.xy_scale, w.wh_scale, loc_weight=2.0)
        loc_loss.backward()
        np.testing.assert_allclose(float(loc_loss), z[f"giou_loss_values_{g}"][2], rtol=1e-6)
        np.testing.assert_allclose(locs.grad.numpy(), z[f"giou_loss_grad_locs_{g}"], rtol=1e-5, atol=1e-8)
