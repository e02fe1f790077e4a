"""GPU parity: the CUDA path (through torch.ops.ssd_b200.* -> C ABI) against the CPU oracle and
against the golden fixtures the reference itself produced.

Bars (BASELINE.json north_star): matched indices, mining selections and NMS keep lists bit
exact on identical fp32 stage inputs; encoded / decoded boxes and scores within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import anchor_pipeline_oracle as ora
from single_shot_detection_b200 import workloads as wl

pytestmark = pytest.mark.gpu

REL = 1e-5          # tolerance of the floating-point stages (north_star: 1e-5 relative)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


@pytest.fixture(scope="module", params=gio.PIPELINE_CASES)
def case(request):
    return gio.PipelineCase(request.param)


def _modules():
    from single_shot_detection_b200 import box_coder, box_utils, matcher, postprocessor, sampler, target_assigner
    return target_assigner, matcher, box_coder, sampler, postprocessor, box_utils


# ----------------------------------------------------------------------------------------------
# target assignment (a1-a4)
# ----------------------------------------------------------------------------------------------
def test_assign_targets_bit_exact_vs_golden(case, dev):
    ta, *_ = _modules()
    assigner = ta.TargetAssigner(case.matched, case.unmatched, nan_check="sync")
    target = assigner.encode_ground_truth(case.gt, case.anchors)
    assert target.is_cuda and target.shape == (case.B, case.A, 6)
    match = assigner.last_match.cpu().long()
    bad = (match != case.match_idx).nonzero()
    assert bad.numel() == 0, f"matched indices differ at {bad[:8].tolist()}"
    assert torch.equal(target.cpu(), case.target)
    stats = assigner.last_stats.cpu()
    cls = case.target[..., 4]
    assert stats[:, 0].tolist() == ((cls != 0) & (cls != -1)).sum(1).tolist()
    assert stats[:, 1].tolist() == (cls == -1).sum(1).tolist()
    assert stats[:, 3].tolist() == [g.shape[0] for g in case.gt]


def test_assign_targets_random_vs_oracle(dev):
    ta, *_ = _modules()
    gen = torch.Generator().manual_seed(7)
    for name, batch, thr in [("tiny_voc_b3", 5, (0.5, 0.5)), ("ssd_mb2_coco_b64", 9, (0.5, 0.4)),
                             ("retina500_coco_b32", 3, (0.5, 0.4)), ("ssd300_voc_8108_b8", 40, (0.6, 0.3))]:
        w = wl.WORKLOADS[name]
        anchors = wl.build_anchors(w)
        gt = wl.make_ground_truth(batch, w.img, w.num_fg, w.max_gt, gen, mixup=0.4 if batch % 2 else None)
        gt[batch // 2] = torch.zeros((0, 6))
        assigner = ta.TargetAssigner(*thr)
        target = assigner.encode_ground_truth(gt, anchors)
        ref, ref_match = ora.assign_targets(gt, anchors, *thr, return_match=True)
        assert torch.equal(assigner.last_match.cpu().long(), torch.stack(ref_match)), name
        assert torch.equal(target.cpu(), ref), name


def test_assign_many_boxes_and_collisions(dev):
    """G larger than a warp, duplicated boxes (ties) and GTs that all force the same anchor."""
    ta, *_ = _modules()
    w = wl.WORKLOADS["tiny_voc_b3"]
    anchors = wl.build_anchors(w)
    gen = torch.Generator().manual_seed(11)
    base = wl.make_ground_truth(1, w.img, w.num_fg, 5, gen)[0]
    many = torch.cat([base] * 40, dim=0)                      # 40 copies -> exact IoU ties
    tiny = torch.tensor([[1., 1., 1.5, 1.5, 2., 1.], [1., 1., 1.25, 1.5, 3., 1.], [1., 1., 1.5, 1.25, 4., 1.]])
    gt = [many, tiny, torch.cat([tiny, base])]
    assigner = ta.TargetAssigner(0.5, 0.2)
    target = assigner.encode_ground_truth(gt, anchors)
    ref, ref_match = ora.assign_targets(gt, anchors, 0.5, 0.2, return_match=True)
    assert torch.equal(assigner.last_match.cpu().long(), torch.stack(ref_match))
    assert torch.equal(target.cpu(), ref)


def test_assign_nan_assert(dev):
    ta, *_ = _modules()
    w = wl.WORKLOADS["tiny_voc_b3"]
    anchors = wl.build_anchors(w)
    gt = [torch.tensor([[float("nan"), 2., 30., 30., 1., 1.]])]
    with pytest.raises(AssertionError):
        ta.TargetAssigner(0.5, 0.5, nan_check="sync").encode_ground_truth(gt, anchors)
    with pytest.raises(AssertionError):
        ta.TargetAssigner(0.3, 0.5).encode_ground_truth([torch.zeros((0, 6))], anchors)


def test_iou_and_matcher_api(case, dev):
    _, matcher, _, _, _, box_utils = _modules()
    if case.gt[0].shape[0] == 0:
        pytest.skip("empty image")
    corners = box_utils.to_corners(case.anchors.to(dev))
    assert torch.equal(corners.cpu(), ora.corners_from_centroids(case.anchors))
    iou = box_utils.iou(case.gt[0][:, :4].contiguous().to(dev), corners)
    ref = ora.pairwise_iou(case.gt[0][:, :4], ora.corners_from_centroids(case.anchors))
    assert torch.equal(iou.cpu(), ref)
    idx = matcher.match_per_prediction(iou, case.matched, case.unmatched)
    assert idx.dtype == torch.int64
    assert torch.equal(idx.cpu(), case.match_idx[0])
    assert matcher.NOT_MATCHED == -2 and matcher.IGNORE == -1


# ----------------------------------------------------------------------------------------------
# box coding (a6, a7)
# ----------------------------------------------------------------------------------------------
def _close(a, b):
    torch.testing.assert_close(a, b, rtol=REL, atol=1e-6, equal_nan=True)


def test_box_coder_all_orders(case, dev):
    _, _, bc, _, _, box_utils = _modules()
    w = case.w
    coder = bc.BoxCoder(w.xy_scale, w.wh_scale, w.eps)
    # the loss route: in place on the strided view target[..., 0:4]          multibox_loss.py:81-82
    t = case.target.to(dev)
    tl = t[..., 0:4]
    assert box_utils.to_centroids(tl, inplace=True) is None
    ref_t = case.target.clone()
    ref_tl = ref_t[..., 0:4]
    ora.centroids_from_corners(ref_tl, inplace=True)
    assert torch.equal(tl.cpu(), ref_tl)
    assert coder.encode_box(tl, case.anchors, inplace=True) is tl
    _close(case.enc_view(tl.cpu()), case.enc_inplace)
    assert torch.equal(case.enc_view(tl.cpu())[..., :2], case.enc_inplace[..., :2])            # xy: no transcendental
    assert torch.equal(t.cpu()[..., 4:], case.target[..., 4:])                   # class/score untouched
    # fused single pass, same rounding
    t2 = case.target.to(dev)
    coder.encode_corners_(t2[..., 0:4], case.anchors)
    assert torch.equal(t2, t)
    # out-of-place orders
    cen = box_utils.to_centroids(case.target[..., 0:4].to(dev))
    ref_cen = ora.centroids_from_corners(case.target[..., 0:4])
    assert torch.equal(cen.cpu(), ref_cen)
    enc = coder.encode_box(cen, case.anchors)
    _close(enc.cpu(), ora.encode_boxes(ref_cen, case.anchors, w.xy_scale, w.wh_scale, w.eps))
    locs = case.locs.view(case.B, case.A, 4)
    dec = coder.decode_box(locs.to(dev), case.anchors)
    ref_dec = ora.decode_boxes(locs, case.anchors, w.xy_scale, w.wh_scale)
    _close(dec.cpu(), ref_dec)
    assert torch.equal(dec.cpu()[..., :2], ref_dec[..., :2])
    dec_in = locs.to(dev).clone()
    coder.decode_box(dec_in, case.anchors, inplace=torch.tensor(1))
    _close(dec_in.cpu(), ora.decode_boxes(locs.clone(), case.anchors, w.xy_scale, w.wh_scale, inplace=True))
    if case.full:
        _close(dec.cpu(), case.t("decoded"))
        _close(enc.cpu(), case.t("enc_oop"))


# ----------------------------------------------------------------------------------------------
# samplers (a5)
# ----------------------------------------------------------------------------------------------
def test_naive_sampler(case, dev):
    _, _, _, sampler, _, _ = _modules()
    cls = case.target[..., 4].long().to(dev)
    mask = sampler.naive_sampler(None, cls)
    assert mask.dtype == torch.bool and torch.equal(mask.cpu(), case.naive_mask)


def test_hard_negative_selection_bit_exact_on_identical_loss(case, dev):
    """Stage boundary: feed the oracle's own fp32 criterion, demand the identical selection."""
    _, _, _, sampler, _, _ = _modules()
    w = case.w
    cls = case.target[..., 4].long()
    loss = ora.background_loss(case.scores.view(case.B, case.A, case.C))
    ref = ora.mine_hard_negatives(None, cls, w.ratio, w.min_neg, canonical=True, loss=loss)
    mask = sampler.hard_negative_mining_from_loss(loss.to(dev), cls.to(dev), w.ratio, w.min_neg)
    assert torch.equal(mask.cpu(), ref)
    tied = ora.mining_boundary_tie(loss, cls, w.ratio, w.min_neg)
    for i in range(case.B):                      # and it is the reference's own answer when unique
        if not tied[i]:
            assert torch.equal(mask[i].cpu(), case.hnm_mask[i])


def test_hard_negative_mining_end_to_end(case, dev):
    _, _, _, sampler, _, _ = _modules()
    w = case.w
    cls = case.target[..., 4].long().to(dev)
    logits = case.scores.view(case.B, case.A, case.C).to(dev)
    mask = sampler.hard_negative_mining(logits, cls, w.ratio, w.min_neg)
    stats = sampler.hard_negative_mining.last_stats.cpu()
    # counts are exact; the set may differ only through ulp-level differences of log_softmax
    assert mask.sum(1).tolist() == case.hnm_mask.sum(1).tolist()
    mism = int((mask.cpu() != case.hnm_mask).sum())
    assert mism <= 2 * case.B, f"{mism} anchors differ from the reference selection"
    cls_cpu = cls.cpu()
    assert stats[:, 0].tolist() == ((cls_cpu != 0) & (cls_cpu != -1)).sum(1).tolist()
    assert stats[:, 1].tolist() == (cls_cpu == 0).sum(1).tolist()


@pytest.mark.parametrize("ratio,min_neg", [(3, 5), (3.0, 5), (2.5, 0), (1, 100000), (0, 0), (7, 1)])
def test_hard_negative_ratio_variants_and_ties(dev, ratio, min_neg):
    """Quantised losses force ties across the cut; the kernel's rule is lower anchor first."""
    _, _, _, sampler, _, _ = _modules()
    gen = torch.Generator().manual_seed(3)
    b, a = 4, 3000
    cls = torch.zeros((b, a), dtype=torch.long)
    cls[torch.rand((b, a), generator=gen) < 0.02] = 3
    cls[torch.rand((b, a), generator=gen) < 0.05] = -1
    cls[3] = 0                                                          # no positives at all
    loss = (torch.rand((b, a), generator=gen) * 8).round() / 8          # heavy ties
    loss[2] = 1.0                                                       # all equal
    ref = ora.mine_hard_negatives(None, cls, ratio, min_neg, canonical=True, loss=loss)
    mask = sampler.hard_negative_mining_from_loss(loss.to(dev), cls.to(dev), ratio, min_neg)
    assert torch.equal(mask.cpu(), ref)


def test_hard_negative_large_anchor_counts(dev):
    """A > 12288 (shared-memory keys) and A > 56000 (global keys) paths of the selection kernel."""
    _, _, _, sampler, _, _ = _modules()
    gen = torch.Generator().manual_seed(5)
    for a in (20000, 70000):
        cls = torch.zeros((2, a), dtype=torch.long)
        cls[torch.rand((2, a), generator=gen) < 0.01] = 1
        logits = torch.randn((2, a, 5), generator=gen)
        loss = ora.background_loss(logits)
        ref = ora.mine_hard_negatives(None, cls, 3, 5, canonical=True, loss=loss)
        mask = sampler.hard_negative_mining_from_loss(loss.to(dev), cls.to(dev), 3, 5)
        assert torch.equal(mask.cpu(), ref), a
        full = sampler.hard_negative_mining(logits.to(dev), cls.to(dev), 3, 5)
        assert int((full.cpu() != ref).sum()) <= 4, a


def test_mining_loss_matches_log_softmax(dev):
    """The streamed criterion itself, read back through the keys: relative 1e-5."""
    _, _, _, sampler, _, _ = _modules()
    gen = torch.Generator().manual_seed(9)
    for c in (2, 6, 21, 32, 64, 81, 91, 200, 300):
        a = 777
        logits = torch.randn((2, a, c), generator=gen) * 3
        cls = torch.zeros((2, a), dtype=torch.long)
        # with ratio huge every negative is selected; with one positive and ratio 1, exactly the
        # single largest loss is selected -> its argmax must match the oracle's
        cls[:, 0] = 1
        mask = sampler.hard_negative_mining(logits.to(dev), cls.to(dev), 1, 0).cpu()
        loss = ora.background_loss(logits)
        loss[:, 0] = -1
        for i in range(2):
            picked = mask[i].nonzero().view(-1).tolist()
            assert len(picked) == 2 and picked[0] == 0
            best = float(loss[i].max())
            assert abs(float(loss[i, picked[1]]) - best) <= REL * abs(best), (c, i)


# ----------------------------------------------------------------------------------------------
# post-processor (a7-a9)
# ----------------------------------------------------------------------------------------------
def _postprocessor(case, max_total):
    _, _, bc, _, pp, _ = _modules()
    w = case.w
    coder = bc.BoxCoder(w.xy_scale, w.wh_scale, w.eps)
    return pp.Postprocessor(coder, w.score_threshold,
                            {"max_per_class": w.max_per_class, "overlap_threshold": w.overlap_threshold},
                            score_converter=w.converter, max_total=max_total)


def _dets_close(m, r, img):
    """Scores within 1e-5 relative.  Corner coordinates are differences of decoded centre/size
    values that are each within 1e-5 relative, so their ABSOLUTE error scales with the image
    size: atol = 2e-5 * img."""
    return (m.shape == r.shape and torch.equal(m[:, 4], r[:, 4])
            and torch.allclose(m[:, 5], r[:, 5], rtol=REL, atol=1e-7)
            and torch.allclose(m[:, :4], r[:, :4], rtol=REL, atol=2e-5 * img))


def _compare_dets(mine, ref, img):
    assert len(mine) == len(ref)
    for i, (m, r) in enumerate(zip(mine, ref)):
        m = m.cpu()
        assert m.shape == r.shape, (i, m.shape, r.shape)
        assert torch.equal(m[:, 4], r[:, 4]), f"image {i}: class column differs"
        torch.testing.assert_close(m[:, 5], r[:, 5], rtol=REL, atol=1e-7)
        torch.testing.assert_close(m[:, :4], r[:, :4], rtol=REL, atol=2e-5 * img)


@pytest.mark.parametrize("max_total", ["cfg", None])
def test_postprocess_vs_reference_golden(case, dev, max_total):
    w = case.w
    mt = w.max_total if max_total == "cfg" else None
    ref = case.dets if max_total == "cfg" else case.dets_all
    post = _postprocessor(case, mt)
    dets = post.postprocess((case.scores.to(dev), case.locs.to(dev)), case.anchors)
    assert post.last_status[0] == 0
    if mt is not None:
        # final top-k: equal scores may be ordered differently by the reference's topk
        for m, r in zip(dets, ref):
            assert m.shape == r.shape
            torch.testing.assert_close(m.cpu()[:, 5], r[:, 5], rtol=REL, atol=1e-6)
        dets = [d.cpu() for d in dets]           # same row order as the reference: sorted only if n > T
        same = all(_dets_close(d, r, w.img) for d, r in zip(dets, ref))
        if not same:
            for d, r in zip(dets, ref):
                assert len(torch.unique(r[:, 5])) < r.shape[0], "rows differ without a score tie"
        return
    _compare_dets(dets, ref, w.img)
    if case.det_all_anchor is not None:
        # BASELINE-config fixtures: the kept rows sit on the anchors the REFERENCE kept, class by class and in its
        # order (from logits the scores carry ulp-level differences, so two nearly equal scores may swap places:
        # the kept SET per class must be identical, the order may differ in a handful of positions)
        counts = [int(d.shape[0]) for d in dets]
        mine = post.last_anchors.cpu()
        swapped = 0
        for i in range(case.B):
            got, want = mine[i, :counts[i]].long(), case.det_all_anchor[i].long()
            cls = ref[i][:, 4].long()
            assert torch.equal(torch.sort(got * 128 + cls)[0], torch.sort(want * 128 + cls)[0]), i
            swapped += int((got != want).sum())
        assert swapped <= 2 * case.B, f"{swapped} kept rows are ordered differently from the reference"


def test_postprocess_stage_exact_keep_lists(case, dev):
    """Identical fp32 stage inputs (the oracle's probabilities and decoded corners) -> the kept
    anchors per class, their order and every output bit must match."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.ops import OPS
    w = case.w
    logits = case.scores.view(case.B, case.A, case.C)
    probs = torch.softmax(logits, -1) if w.converter == "SOFTMAX" else torch.sigmoid(logits)
    first_fg = 1 if w.converter == "SOFTMAX" else 0
    fg = probs[..., first_fg:]
    corners = ora.corners_from_centroids(ora.decode_boxes(case.locs.view(case.B, case.A, 4), case.anchors,
                                                          w.xy_scale, w.wh_scale))
    assert not bool(ora.class_topk_boundary_tie(fg, w.score_threshold, w.max_per_class).any())
    ref, ref_keep = ora.detections_from_scores(fg, corners, w.score_threshold, w.overlap_threshold,
                                               w.max_per_class, None, canonical=True, return_keep=True)
    dets, counts, anchors, status = OPS.postprocess(probs.contiguous().to(dev), corners.contiguous().to(dev), None,
                                                    N.CONVERT_IDENTITY, first_fg, N.BOXES_CORNERS, 1.0, 1.0,
                                                    float(w.score_threshold), w.max_per_class,
                                                    float(w.overlap_threshold), 0)
    assert status.tolist()[0] == 0
    counts = counts.tolist()
    for i in range(case.B):
        assert counts[i] == ref[i].shape[0], i
        assert torch.equal(dets[i, :counts[i]].cpu(), ref[i]), i
        assert anchors[i, :counts[i]].cpu().tolist() == torch.cat(ref_keep[i]).tolist(), i
        if case.det_all_anchor is not None:          # ... which are the anchors the REFERENCE kept (make_golden.py)
            assert anchors[i, :counts[i]].cpu().tolist() == case.det_all_anchor[i].tolist(), i
    # with the final top-k: descending score, ties by class-major position
    ref_t = ora.detections_from_scores(fg, corners, w.score_threshold, w.overlap_threshold, w.max_per_class,
                                       w.max_total, canonical=True)
    dets, counts, anchors, status = OPS.postprocess(probs.contiguous().to(dev), corners.contiguous().to(dev), None,
                                                    N.CONVERT_IDENTITY, first_fg, N.BOXES_CORNERS, 1.0, 1.0,
                                                    float(w.score_threshold), w.max_per_class,
                                                    float(w.overlap_threshold), w.max_total)
    for i in range(case.B):
        assert torch.equal(dets[i, :counts[i]].cpu(), ref_t[i]), i


def test_postprocess_sparse_and_empty_classes(dev):
    """Realistic score sparsity: most classes have no candidate, some have fewer than K."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.ops import OPS
    gen = torch.Generator().manual_seed(21)
    b, a, c = 3, 2268, 21
    probs = torch.rand((b, a, c), generator=gen) * 0.009
    hot = torch.rand((b, a, c), generator=gen) < 0.002
    probs[hot] = torch.rand((int(hot.sum()),), generator=gen)
    probs[1] = 0.0                                                     # an image with no detection
    cxy = torch.rand((b, a, 2), generator=gen) * 300
    wh = torch.rand((b, a, 2), generator=gen) * 80 + 2
    corners = torch.cat([cxy - wh / 2, cxy + wh / 2], dim=-1)
    ref, ref_keep = ora.detections_from_scores(probs[..., 1:], corners, 0.01, 0.45, 100, 200, canonical=True,
                                               return_keep=True)
    dets, counts, anchors, status = OPS.postprocess(probs.to(dev), corners.to(dev), None, N.CONVERT_IDENTITY, 1,
                                                    N.BOXES_CORNERS, 1.0, 1.0, 0.01, 100, 0.45, 200)
    assert status.tolist()[0] == 0
    for i in range(b):
        assert int(counts[i]) == ref[i].shape[0]
        assert torch.equal(dets[i, :int(counts[i])].cpu(), ref[i])


def test_postprocess_score_ties_and_duplicates(dev):
    """Exact score ties inside a class (lower anchor first) and identical boxes (suppressed)."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.ops import OPS
    gen = torch.Generator().manual_seed(22)
    b, a, c = 2, 1500, 3
    probs = (torch.rand((b, a, c), generator=gen) * 64).round() / 64          # 65 distinct values
    cxy = (torch.rand((b, a, 2), generator=gen) * 20).round() * 10
    wh = (torch.rand((b, a, 2), generator=gen) * 4).round() * 8 + 8
    corners = torch.cat([cxy - wh / 2, cxy + wh / 2], dim=-1)
    ref = ora.detections_from_scores(probs, corners, 0.3, 0.5, 50, None, canonical=True)
    dets, counts, anchors, status = OPS.postprocess(probs.to(dev), corners.to(dev), None, N.CONVERT_IDENTITY, 0,
                                                    N.BOXES_CORNERS, 1.0, 1.0, 0.3, 50, 0.5, 0)
    for i in range(b):
        assert int(counts[i]) == ref[i].shape[0]
        assert torch.equal(dets[i, :int(counts[i])].cpu(), ref[i])


def test_postprocess_overflow_fallback_is_exact(dev):
    """Dense / fully tied scores overflow the candidate lists; the column-rescan fallback must
    still give the canonical answer (ties -> lower anchor) for every converter."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.ops import OPS
    gen = torch.Generator().manual_seed(31)
    b, a, c = 2, 5000, 4
    cxy = torch.rand((b, a, 2), generator=gen) * 300
    wh = torch.rand((b, a, 2), generator=gen) * 30 + 2
    corners = torch.cat([cxy - wh / 2, cxy + wh / 2], dim=-1)
    probs = torch.full((b, a, c), 0.5)
    probs[0, :, 1] = (torch.rand((a,), generator=gen) * 4).round() / 8 + 0.25       # 5 distinct values
    probs[1, :, 2] = torch.linspace(0.02, 0.9, a)                                    # sorted ascending
    probs[1, :, 3] = torch.linspace(0.9, 0.02, a)                                    # sorted descending
    ref = ora.detections_from_scores(probs, corners, 0.01, 0.45, 100, 200, canonical=True)
    dets, counts, anchors, status = OPS.postprocess(probs.to(dev), corners.to(dev), None, N.CONVERT_IDENTITY, 0,
                                                    N.BOXES_CORNERS, 1.0, 1.0, 0.01, 100, 0.45, 200)
    st = status.tolist()
    assert st[0] == 0 and st[1] > 0, st           # the fallback really ran
    for i in range(b):
        assert int(counts[i]) == ref[i].shape[0]
        assert torch.equal(dets[i, :int(counts[i])].cpu(), ref[i])
    # saturated logits: sigmoid == 1.0f for many anchors, softmax one-hot
    logits = torch.full((1, a, c), -20.0)
    logits[0, ::3, 1] = 30.0
    logits[0, 1::3, 2] = 25.0 + torch.rand((len(range(1, a, 3)),), generator=gen)
    for conv, name, first in ((N.CONVERT_SIGMOID, "SIGMOID", 0), (N.CONVERT_SOFTMAX, "SOFTMAX", 1)):
        fg = ora.convert_scores(logits, name)
        ref = ora.detections_from_scores(fg, corners[:1], 0.01, 0.45, 100, 200, canonical=True)
        dets, counts, anchors, status = OPS.postprocess(logits.to(dev), corners[:1].contiguous().to(dev), None, conv,
                                                        first, N.BOXES_CORNERS, 1.0, 1.0, 0.01, 100, 0.45, 200)
        n = int(counts[0])
        assert n == ref[0].shape[0], name
        torch.testing.assert_close(dets[0, :n].cpu(), ref[0], rtol=REL, atol=1e-6)


@pytest.mark.parametrize("conv,c,a", [("SOFTMAX", 21, 8732), ("SOFTMAX", 7, 6500), ("SOFTMAX", 31, 6700), ("SOFTMAX", 19, 7001),
                                      ("SIGMOID", 5, 6401), ("SIGMOID", 27, 9000), ("SOFTMAX", 21, 6399), ("SIGMOID", 20, 7000)])
def test_postprocess_row_per_lane_shapes_vs_oracle(dev, conv, c, a):
    """The row-per-lane shapes of the streaming kernels (C <= 8 and odd C <= 32): with >= 6400 anchors a warp step of
    32 rows is one block of pass 1 (the block-per-step instantiation), partial last tiles included; 6399 anchors and
    the even C take the generic block bookkeeping.  Logits -> detections against the oracle (keep lists exact through
    the anchors, scores / boxes to REL)."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.ops import OPS
    gen = torch.Generator().manual_seed(1000 + c + a)
    b = 3
    logits = torch.randn((b, a, c), generator=gen) + (-3.0 if conv == "SIGMOID" else 0.0)
    cxy = torch.rand((b, a, 2), generator=gen) * 300
    wh = torch.rand((b, a, 2), generator=gen) * 60 + 2
    corners = torch.cat([cxy - wh / 2, cxy + wh / 2], dim=-1)
    fg = ora.convert_scores(logits, conv)
    # (no final top-k: among the 200 best of ~2000 kept rows two scores 1e-7 apart may swap between the host's softmax
    # and the device's; class-major lists only depend on the order inside a class)
    ref = ora.detections_from_scores(fg, corners, 0.01, 0.45, 100, None, canonical=True)
    code, first = (N.CONVERT_SOFTMAX, 1) if conv == "SOFTMAX" else (N.CONVERT_SIGMOID, 0)
    dets, counts, anchors, status = OPS.postprocess(logits.to(dev), corners.to(dev), None, code, first, N.BOXES_CORNERS,
                                                    1.0, 1.0, 0.01, 100, 0.45, 0)
    assert status.tolist()[1] == 0, status.tolist()          # no list overflowed: the gate path itself is what ran
    _compare_dets([dets[i, :int(counts[i])] for i in range(b)], ref, 300)


def test_nms_api_vs_torchvision_golden(dev):
    _, _, _, _, _, box_utils = _modules()
    z = gio.load("nms.npz")
    for i in range(int(z["num_cases"])):
        bx, sc = torch.from_numpy(z[f"boxes_{i}"]), torch.from_numpy(z[f"scores_{i}"])
        if bx.shape[0] > 512:
            continue
        (bk, sk), keep = box_utils.nms(bx.to(dev), sc.to(dev), float(z[f"thr_{i}"]), 0.01, None)
        assert keep.cpu().tolist() == z[f"keep_{i}"].tolist(), i
    for i in range(int(z["num_topk_cases"])):
        bx, sc = torch.from_numpy(z[f"topk_boxes_{i}"]), torch.from_numpy(z[f"topk_scores_{i}"])
        (bk, sk), keep = box_utils.nms(bx.to(dev), sc.to(dev), 0.45, 0.01, int(z[f"topk_k_{i}"]))
        assert torch.equal(bk.cpu(), torch.from_numpy(z[f"topk_kept_boxes_{i}"])), i
        assert torch.equal(sk.cpu(), torch.from_numpy(z[f"topk_kept_scores_{i}"])), i


def test_nms_threshold_boundaries_exact(dev):
    """The overlap test is `(double)fp32(inter/union) > threshold`.  The kernel evaluates it without a
    division (exact product in double against the rounding midpoint): thresholds placed exactly on,
    one double-ulp below / above and one float-ulp around achieved IoU values must give torchvision's
    keep list, as must negative, zero and > 1 thresholds (division path)."""
    import torchvision
    _, _, _, _, _, box_utils = _modules()
    gen = torch.Generator().manual_seed(77)
    n = 160
    xy = torch.randint(0, 40, (n, 2), generator=gen).float()
    wh = torch.randint(1, 24, (n, 2), generator=gen).float()
    boxes = torch.cat([xy, xy + wh], dim=1)
    boxes[5] = boxes[4]                                    # duplicates (IoU exactly 1)
    boxes[9, 2:] = boxes[9, :2]                            # zero-area box
    scores = torch.rand((n,), generator=gen)
    iou = torchvision.ops.box_iou(boxes, boxes)
    vals = torch.unique(iou[(iou > 0.05) & (iou < 0.95)])
    picks = vals[torch.linspace(0, vals.numel() - 1, 12).long()].tolist()
    thresholds = [0.0, -0.25, 1.0, 1.5, 0.45, 0.5, 1e-35]
    for v in picks:
        v32 = np.float32(v)
        d = float(v32)
        thresholds += [d, float(np.nextafter(d, 0.0)), float(np.nextafter(d, 1.0)),
                       float(np.nextafter(v32, np.float32(0))), float(np.nextafter(v32, np.float32(1)))]
    bd, sd = boxes.to(dev), scores.to(dev)
    for thr in thresholds:
        ref = torchvision.ops.nms(boxes, scores, thr).tolist()
        assert ora.greedy_nms(boxes.numpy(), scores.numpy(), thr).tolist() == ref, thr
        (_, _), keep = box_utils.nms(bd, sd, thr, 0.01, None)
        assert keep.cpu().tolist() == ref, thr


def test_postprocessor_interface(dev):
    _, _, bc, _, pp, _ = _modules()
    coder = bc.BoxCoder(10.0, 5.0)
    with pytest.raises(ValueError):
        pp.Postprocessor(coder, 0.01, {"max_per_class": 100, "overlap_threshold": 0.45}, score_converter="TANH")
    post = pp.Postprocessor(coder, 0.01, {"max_per_class": 100, "overlap_threshold": 0.45}, max_total=200)
    assert post.box_coder is coder and post.score_threshold == 0.01 and post.max_total == 200
    with pytest.raises(TypeError):
        post.postprocess((torch.zeros(1, 40), torch.zeros(1, 16)), torch.ones(4, 4))


# ----------------------------------------------------------------------------------------------
# whole timed region, BASELINE-sized shapes: properties that do not need the (slow) oracle
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,batch", [("ssd300_voc_b32", 32), ("ssd512_coco_b32", 4), ("retina500_coco_b32", 2)])
def test_full_size_properties(dev, name, batch):
    from single_shot_detection_b200.pipeline import AnchorPipeline
    w = wl.WORKLOADS[name]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=23, batch=batch)
    pipe = AnchorPipeline(w.cfg())
    target, mask, dets = pipe.step(gt, anchors, scores, locs)
    a = anchors.shape[0]
    match = pipe.target_assigner.last_match.cpu()
    cls = target[..., 4].cpu()
    # every ground-truth box owns at least one anchor unless a later box took its only one
    for i, g in enumerate(gt):
        owned = set(match[i][match[i] >= 0].tolist())
        assert len(owned) >= 1 and max(owned) < g.shape[0]
    # mining: positives always sampled, ignored never, count = pos + min(max(3*pos,5), neg)
    m = mask.cpu()
    pos = (cls != 0) & (cls != -1)
    assert bool((m & pos).eq(pos).all()) and not bool((m & (cls == -1)).any())
    if w.sampler == "hard_negative_mining":
        n_pos, n_neg = pos.sum(1), (cls == 0).sum(1)
        want = n_pos + torch.minimum(torch.clamp(n_pos * w.ratio, min=w.min_neg), n_neg)
        assert m.sum(1).tolist() == want.tolist()
    else:
        assert torch.equal(m, pos)
    # detections: at most T rows, scores above threshold, sorted descending when the top-k hit,
    # classes in range, boxes well formed, and idempotent (same call, same bits)
    dets2 = pipe.postprocessor.postprocess((scores.to(dev), locs.to(dev)), anchors)
    for d, d2 in zip(dets, dets2):
        d = d.cpu()
        assert torch.equal(d, d2.cpu())
        assert d.shape[0] <= w.max_total and d.shape[1] == 6
        assert bool((d[:, 5] > w.score_threshold).all())
        if d.shape[0] == w.max_total:
            assert bool((d[1:, 5] <= d[:-1, 5]).all())
        assert bool(((d[:, 4] >= 1) & (d[:, 4] <= w.num_fg)).all())
        assert bool((d[:, 2] >= d[:, 0]).all()) and bool((d[:, 3] >= d[:, 1]).all())


def test_full_size_ssd300_b8_vs_oracle(dev):
    """BASELINE configs[0] at full size against the oracle (a few seconds of CPU)."""
    from single_shot_detection_b200.pipeline import AnchorPipeline
    w = wl.WORKLOADS["ssd300_voc_b8"]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=23)
    pipe = AnchorPipeline(w.cfg())
    target, mask, dets = pipe.step(gt, anchors, scores, locs)
    ref_target, ref_mask, ref_dets = ora.run_step(gt, anchors, scores, locs, w.cfg(), canonical=True,
                                                   use_torchvision=False)
    pos = ora.positives_mask(target[..., 4].cpu().long())
    assert torch.equal(target[..., 4:].cpu(), ref_target[..., 4:])
    torch.testing.assert_close(target[..., :4].cpu()[pos], ref_target[..., :4][pos], rtol=REL, atol=1e-6)
    assert int((mask.cpu() != ref_mask).sum()) <= 8
    _compare_dets(dets, ref_dets, w.img)


def test_stream_matches_step(dev):
    """The software-pipelined form (H2D of batch i+1 under the kernels of batch i, detections read
    back to pinned host memory) returns exactly what step() returns, batch by batch, in order."""
    from single_shot_detection_b200.pipeline import AnchorPipeline
    w = wl.WORKLOADS["ssd300_voc_b8"]
    batches = []
    for s in range(5):
        anchors, gt, scores, locs = wl.make_inputs(w, seed=100 + s, batch=3 + (s % 2))
        batches.append((gt, scores.pin_memory(), locs.pin_memory()))
    ref_pipe = AnchorPipeline(w.cfg())
    expected = []
    for gt, scores, locs in batches:
        target, mask, dets = ref_pipe.step(gt, anchors, scores, locs)
        expected.append((target.cpu(), mask.cpu(), [d.cpu() for d in dets]))
    pipe = AnchorPipeline(w.cfg())
    seen = 0
    for (target, mask, dets), (rt, rm, rd) in zip(pipe.stream(iter(batches), anchors), expected):
        assert torch.equal(target.cpu(), rt) and torch.equal(mask.cpu(), rm)
        assert len(dets) == len(rd)
        for d, r in zip(dets, rd):
            assert not d.is_cuda and torch.equal(d, r)
        seen += 1
    assert seen == len(batches)


# ----------------------------------------------------------------------------------------------
# masked multibox losses (a10 / f1): fused forward + gradient
# ----------------------------------------------------------------------------------------------
def _loss_module(w):
    import functools
    from single_shot_detection_b200 import box_coder, multibox_loss, sampler
    coder = box_coder.BoxCoder(w.xy_scale, w.wh_scale, w.eps)
    if w.converter == "SOFTMAX":
        smp = functools.partial(sampler.hard_negative_mining, negative_per_positive_ratio=w.ratio,
                                min_negative_per_image=w.min_neg)
        return multibox_loss.MultiboxLoss(smp, coder, {"name": "CrossEntropyLoss"}, {"name": "SmoothL1Loss"})
    return multibox_loss.MultiboxLoss(sampler.naive_sampler, coder,
                                      {"name": "SigmoidFocalLoss", "gamma": 2.0, "alpha": 0.25},
                                      {"name": "SmoothL1Loss"})


def test_multibox_loss_vs_reference_golden(case, dev):
    """(loss, class_loss, loc_loss) of the reference's own MultiboxLoss on the fixtures, 1e-5 relative."""
    crit = _loss_module(case.w)
    target = case.target.clone().to(dev)
    loss3 = crit((case.scores.to(dev), case.locs.to(dev)), case.anchors, target)
    got = [float(x) for x in loss3]
    np.testing.assert_allclose(got, case.loss3, rtol=REL)
    # the box columns of `target` were encoded in place, as the reference's forward does
    pos = ora.positives_mask(case.target[..., 4].long())
    if case.enc_rows is None:
        torch.testing.assert_close(target[..., :4].cpu()[pos], case.enc_inplace[pos], rtol=REL, atol=1e-6)
    else:                                   # the fixture holds every matched row (and a sample of the others)
        sel = pos.reshape(-1)[case.enc_rows]
        assert int(sel.sum()) == int(pos.sum())
        torch.testing.assert_close(case.enc_view(target[..., :4].cpu())[sel], case.enc_inplace[sel], rtol=REL, atol=1e-6)


@pytest.mark.parametrize("kind", ["ce", "focal"])
def test_multibox_loss_gradients_vs_autograd(dev, kind):
    """Dense gradients of the fused pass against torch autograd through the oracle's loss, with the
    SAME sampler mask and encoded targets; unequal weights and a non-unit upstream gradient."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.multibox_loss import _FusedMultiboxLoss
    w = wl.WORKLOADS["tiny_voc_b3" if kind == "ce" else "tiny_sigmoid_b2"]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=77, batch=4)
    b, a = 4, anchors.shape[0]
    target = ora.assign_targets(gt, anchors, w.matched_threshold, w.unmatched_threshold)
    cls = target[..., 4].long()
    if kind == "ce":
        mask = ora.mine_hard_negatives(scores.view(b, a, -1), cls, 3, 5, canonical=True)
        mask[0, :7] = True                      # a few extra anchors, including ignored ones if any
    else:
        mask = ora.positives_mask(cls)
        mask[1, :9] = True                      # negatives in the mask: all-zero focal targets
        target[..., 5] = torch.where(ora.positives_mask(cls), torch.rand(cls.shape) * 0.5 + 0.5, target[..., 5])
    tl = target[..., :4]
    ora.centroids_from_corners(tl, inplace=True)
    ora.encode_boxes(tl, anchors, w.xy_scale, w.wh_scale, w.eps, inplace=True)
    cw, lw = 1.7, 0.6
    s_ref = scores.clone().requires_grad_(True)
    l_ref = locs.clone().requires_grad_(True)
    if kind == "ce":
        _, c_ref, r_ref = ora.multibox_loss_ce_smoothl1(s_ref, l_ref, anchors, target, mask, tl)
    else:
        _, c_ref, r_ref = ora.multibox_loss_focal_smoothl1(s_ref, l_ref, target, mask, tl)
    ref3 = torch.stack([cw * c_ref + lw * r_ref, cw * c_ref, lw * r_ref])
    up = torch.tensor([0.9, 0.3, -0.2])
    (ref3 * up).sum().backward()

    s_dev = scores.to(dev).requires_grad_(True)
    l_dev = locs.to(dev).requires_grad_(True)
    loss3 = _FusedMultiboxLoss.apply(s_dev, l_dev, target.to(dev), mask.to(dev),
                                     N.LOSS_SOFTMAX_CE if kind == "ce" else N.LOSS_SIGMOID_FOCAL, 2.0, 0.25, cw, lw)
    torch.testing.assert_close(loss3.cpu(), ref3.detach(), rtol=REL, atol=1e-7)
    (loss3 * up.to(dev)).sum().backward()
    scale = float(s_ref.grad.abs().max())
    torch.testing.assert_close(s_dev.grad.cpu(), s_ref.grad, rtol=1e-4, atol=1e-6 * max(scale, 1e-3))
    torch.testing.assert_close(l_dev.grad.cpu(), l_ref.grad, rtol=1e-4, atol=1e-7)
    # rows outside the masks carry exactly zero gradient
    assert float(s_dev.grad.view(b, a, -1)[~mask.to(dev)].abs().max()) == 0.0


# ----------------------------------------------------------------------------------------------
# anchor tables written on the device
# ----------------------------------------------------------------------------------------------
def test_device_anchor_tables_bit_exact_vs_reference_tables(dev):
    from single_shot_detection_b200 import anchor_generators as ag
    z = gio.load("anchors.npz")
    for name in z.files:
        w = wl.WORKLOADS[name]
        gens = wl.build_anchor_generators(w)
        table = ag.generate_anchors(gens, (w.img, w.img), [(s, s) for s in w.fmaps], dev)
        assert table.is_cuda and torch.equal(table.cpu(), torch.from_numpy(z[name])), name
        # per-level API of the reference: [H, W, boxes, 4]
        img = torch.empty((1, 3, w.img, w.img))
        fm = torch.empty((1, 1, w.fmaps[0], w.fmaps[0]), device=dev)
        lvl = gens[0].generate(img, fm)
        n = w.fmaps[0] ** 2 * gens[0].num_boxes
        assert lvl.shape == (w.fmaps[0], w.fmaps[0], gens[0].num_boxes, 4)
        assert torch.equal(lvl.reshape(-1, 4).cpu(), torch.from_numpy(z[name][:n]))


def test_device_linspace_matches_torch_cpu_for_odd_sizes(dev):
    """Cell centres for every map size 1..70 at several image sizes, non-square maps and explicit steps:
    the kernel's fma form of torch.linspace against the CPU kernel of the installed torch."""
    from single_shot_detection_b200 import anchor_generators as ag
    for img_w, img_h in [(300, 300), (512, 384), (500, 333), (321, 1025)]:
        for cells in list(range(1, 71)) + [100, 128, 129]:
            fm = (cells, max(1, (cells * 3) // 4))
            for step in (None, 8):
                gen = ag.SsdAnchorGenerator([1.0, 2.0], min_scale=0.2, max_scale=0.4, step=step)
                got = ag.generate_anchors([gen], (img_w, img_h), [fm], dev, cache=False).cpu().view(fm[1], fm[0], -1, 4)
                sw, sh = (step, step) if step else (img_w / fm[0], img_h / fm[1])
                xs = torch.linspace(0.5 * sw, (0.5 + fm[0] - 1) * sw, fm[0])
                ys = torch.linspace(0.5 * sh, (0.5 + fm[1] - 1) * sh, fm[1])
                assert torch.equal(got[0, :, 0, 0], xs), (img_w, cells, step)
                assert torch.equal(got[:, 0, 0, 1], ys), (img_h, cells, step)


def test_pack_shard_kernel_equals_tensor_op_packing(dev):
    """csrc/exchange.cu against the tensor-op packing the gloo tests exercise on the CPU (sharding.pack_shard)."""
    from single_shot_detection_b200 import sharding
    from single_shot_detection_b200.pipeline import matched_stats
    gen = torch.Generator().manual_seed(11)
    for n, t, cap in [(5, 7, 5), (3, 200, 8), (0, 4, 2), (32, 200, 32)]:
        dets = torch.rand((n, t, 6), generator=gen)
        counts = torch.randint(0, t + 1, (n,), generator=gen, dtype=torch.int32)
        a_stats = torch.randint(0, 99, (n, 4), generator=gen, dtype=torch.int32)
        m_stats = torch.randint(0, 99, (n, 4), generator=gen, dtype=torch.int32)
        want_stats = matched_stats(a_stats, m_stats, counts)
        want = sharding.pack_shard(dets, counts, want_stats, cap)
        got, got_stats = sharding.pack_shard_device(dets.to(dev), counts.to(dev), a_stats.to(dev), m_stats.to(dev), cap)
        assert torch.equal(got.cpu().view(torch.int32), want.view(torch.int32)), (n, t, cap)
        assert torch.equal(got_stats.cpu(), want_stats)
        got2, _ = sharding.pack_shard_device(dets.to(dev), counts.to(dev), a_stats.to(dev), None, cap)
        want2 = sharding.pack_shard(dets, counts, matched_stats(a_stats, None, counts), cap)
        assert torch.equal(got2.cpu().view(torch.int32), want2.view(torch.int32))


# ----------------------------------------------------------------------------------------------
# eval step: the mining criterion out of the post-processor's first pass (one read of the logits)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,batch", [("ssd300_voc_b8", 8), ("tiny_voc_b3", 3), ("ssd_mb2_coco_b64", 6)])
def test_shared_logit_pass_equals_separate_calls(dev, name, batch):
    from single_shot_detection_b200.pipeline import AnchorPipeline
    from single_shot_detection_b200.target_assigner import pack_ground_truth
    w = wl.WORKLOADS[name]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=5, batch=batch)
    outs = []
    for share in (False, True):
        pipe = AnchorPipeline(w.cfg())
        pipe.share_logit_pass = share
        packed = pack_ground_truth(gt, dev)
        out = pipe.step_device(packed, anchors.to(dev), scores.to(dev), locs.to(dev))
        torch.cuda.synchronize()
        outs.append(out)
    a, b = outs
    assert torch.equal(a.target, b.target)
    assert torch.equal(a.dets, b.dets) and torch.equal(a.counts, b.counts)
    assert torch.equal(a.mining_stats[:, :3], b.mining_stats[:, :3])           # positives, negatives, selected
    diff = (a.mask != b.mask)
    if w.num_score_cols == 21 or w.num_score_cols <= 8:
        assert not bool(diff.any())          # same row shape in both kernels: bit-identical criterion
    else:
        # C = 81: pass 1 sums a row over 8 lanes, the sampler's own kernel over 4 -- the criterion can differ
        # in the last bit, which may swap two anchors exactly at the cut
        assert int(diff.sum()) <= 2 * batch
    # and the list API takes the same route
    pipe = AnchorPipeline(w.cfg())
    target, mask, dets = pipe.step(gt, anchors, scores.to(dev), locs.to(dev))
    assert torch.equal(mask, b.mask) and torch.equal(target, b.target)
    for i, d in enumerate(dets):
        assert torch.equal(d, b.dets[i, : int(b.counts[i])])


@pytest.mark.parametrize("name,batch", [("ssd300_voc_b8", 8), ("edge", 0), ("ssd_mb2_coco_b64", 12), ("retina500_coco_b32", 3)])
def test_assignment_with_coded_boxes_equals_the_two_box_passes(dev, name, batch):
    """ssd_assign_targets_encoded == ssd_assign_targets + to_centroids(inplace) + encode_box(inplace), bit for bit,
    including the rows the forced matches rewrite and the statistics."""
    ta, _, bc, _, _, box_utils = _modules()
    if name == "edge":
        case = gio.PipelineCase("edge_tiny")
        w, anchors, gt = case.w, case.anchors, case.gt
    else:
        w = wl.WORKLOADS[name]
        anchors = wl.build_anchors(w)
        gen = torch.Generator().manual_seed(3)
        gt = wl.make_ground_truth(batch, w.img, w.num_fg, w.max_gt, gen, mixup=0.3)
        gt[1] = torch.zeros((0, 6))
        gt[2] = torch.cat([gt[2], gt[2][:1]])                    # duplicate box: colliding forced matches
    coder = bc.BoxCoder(w.xy_scale, w.wh_scale, w.eps)
    from single_shot_detection_b200.target_assigner import pack_ground_truth
    anchors_d = anchors.to(dev)
    a1 = ta.TargetAssigner(w.matched_threshold, w.unmatched_threshold, nan_check="off")
    t1 = a1.encode_packed(pack_ground_truth(gt, dev), anchors_d)
    tl = t1[..., 0:4]
    box_utils.to_centroids(tl, inplace=True)
    coder.encode_box(tl, anchors_d, inplace=True)
    a2 = ta.TargetAssigner(w.matched_threshold, w.unmatched_threshold, nan_check="off")
    t2 = a2.encode_packed(pack_ground_truth(gt, dev), anchors_d, box_coder=coder)
    assert torch.equal(t1.view(torch.int32), t2.view(torch.int32))
    assert torch.equal(a1.last_match, a2.last_match) and torch.equal(a1.last_stats, a2.last_stats)


def test_two_step_graphs_in_flight_equal_serial_steps(dev):
    """bench.py's throughput mode: consecutive steps replayed concurrently on two streams, each slot with its own
    scratch buffers (ops.workspace_slot) -- every step's outputs equal the same step run alone."""
    from single_shot_detection_b200.pipeline import AnchorPipeline
    from single_shot_detection_b200.target_assigner import pack_ground_truth
    w = wl.WORKLOADS["ssd300_voc_b8"]
    anchors = wl.build_anchors(w).to(dev)
    sets, want = [], []
    for k in range(4):
        _, gt, scores, locs = wl.make_inputs(w, seed=40 + k, batch=6)
        packed = pack_ground_truth(gt, dev)
        packed.rows, packed.offsets = packed.rows.clone(), packed.offsets.clone()
        sets.append((packed, scores.to(dev), locs.to(dev)))
        out = AnchorPipeline(w.cfg()).step_device(packed, anchors, sets[-1][1], sets[-1][2])
        torch.cuda.synchronize()
        want.append([t.clone() for t in (out.target, out.mask, out.dets, out.counts)])
    pipes, outs = [], []
    for k, (packed, scores, locs) in enumerate(sets):
        pipe = AnchorPipeline(w.cfg(), workspace_slot=k % 2)
        outs.append(pipe.capture(packed, anchors, scores, locs))
        pipes.append(pipe)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(25):
        for k, pipe in enumerate(pipes):
            with torch.cuda.stream(streams[k % 2]):
                pipe.replay()
    torch.cuda.synchronize()
    for k, out in enumerate(outs):
        got = (out.target, out.mask, out.dets, out.counts)
        for g, r in zip(got, want[k]):
            if g is out.dets:
                for i in range(g.shape[0]):
                    n = int(out.counts[i])
                    assert torch.equal(g[i, :n], r[i, :n]), k
            else:
                assert torch.equal(g, r), k


@pytest.mark.parametrize("chained", [False, True])
def test_step_group_equals_single_steps(dev, chained):
    """pipeline.StepGroup: several steps in ONE graph -- as parallel branches (own workspace slots) or chained one
    after the other (bench.py's serial leg, captured with pass 1 ahead of the assignment branch) -- give the outputs
    of the same steps run alone."""
    from single_shot_detection_b200.pipeline import AnchorPipeline, StepGroup
    from single_shot_detection_b200.target_assigner import pack_ground_truth
    w = wl.WORKLOADS["ssd300_voc_b8"]
    anchors = wl.build_anchors(w).to(dev)
    items, want = [], []
    for k in range(3):
        _, gt, scores, locs = wl.make_inputs(w, seed=60 + k, batch=5)
        packed = pack_ground_truth(gt, dev)
        packed.rows, packed.offsets = packed.rows.clone(), packed.offsets.clone()
        scores, locs = scores.to(dev), locs.to(dev)
        out = AnchorPipeline(w.cfg()).step_device(packed, anchors, scores, locs)
        torch.cuda.synchronize()
        want.append([t.clone() for t in (out.target, out.mask, out.dets, out.counts)])
        pipe = AnchorPipeline(w.cfg(), workspace_slot=0 if chained else k)
        pipe.pass1_first = chained
        items.append((pipe, packed, anchors, scores, locs, {}))
    group = StepGroup(items, chained=chained)
    torch.cuda.synchronize()
    for rep in range(10):
        group.replay()
    torch.cuda.synchronize()
    for k, out in enumerate(group.outs):
        for g, r in zip((out.target, out.mask, out.counts), (want[k][0], want[k][1], want[k][3])):
            assert torch.equal(g, r), k
        for i in range(out.dets.shape[0]):
            n = int(out.counts[i])
            assert torch.equal(out.dets[i, :n], want[k][2][i, :n]), k


def test_peer_exchange_single_rank_equals_tensor_op_packing(dev):
    """sharding.PeerExchange with a world of one (the multi-GPU check is tools/exchange_check.py under torchrun):
    slots, launch counters, row flags, the wait kernel and the gathered layout; no process group needed."""
    from single_shot_detection_b200 import sharding
    from single_shot_detection_b200.pipeline import matched_stats
    for n, t in [(6, 20), (3, 200), (1, 7)]:
        px = sharding.PeerExchange(n, t, slots=2, device=dev)
        assert px.world == 1 and px.gathered(0).shape == (n, sharding.row_words(t)) and sharding.row_words(t) % 4 == 0
        gen = torch.Generator().manual_seed(2)
        for rnd in range(3):
            for k in range(2):
                dets = torch.rand((n, t, 6), generator=gen)
                counts = torch.randint(0, t + 1, (n,), generator=gen, dtype=torch.int32)
                a_stats = torch.randint(0, 9, (n, 4), generator=gen, dtype=torch.int32)
                px.open(k)
                stats = px.pack_exchange(dets.to(dev), counts.to(dev), a_stats.to(dev), None, k)
                px.wait(k)
                want_stats = matched_stats(a_stats, None, counts)
                want = sharding.pack_shard(dets, counts, want_stats, n)
                assert torch.equal(px.gathered(k).cpu().view(torch.int32), want.view(torch.int32)), (rnd, k)
                assert torch.equal(stats.cpu(), want_stats)
        assert px.error() == 0
        px.close()


def test_peer_exchange_in_step_graph_single_rank(dev):
    """The exchange as the last kernel of a captured step graph (what bench.py replays, also at N = 1): the gathered
    slot holds the step's detections, counts and statistics after every replay."""
    from single_shot_detection_b200 import sharding
    from single_shot_detection_b200.pipeline import AnchorPipeline
    from single_shot_detection_b200.target_assigner import pack_ground_truth
    w = wl.WORKLOADS["ssd300_voc_b8"]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=3, batch=4)
    px = sharding.PeerExchange(4, w.max_total, slots=1, device=dev)
    pipe = AnchorPipeline(w.cfg())
    packed = pack_ground_truth(gt, dev)
    out = pipe.capture(packed, anchors.to(dev), scores.to(dev), locs.to(dev), exchange=(px, 0))
    for _ in range(3):
        pipe.replay()
    px.wait(0)
    torch.cuda.synchronize()
    dets, counts, stats = px.unpack(0)
    assert torch.equal(counts.cpu(), out.counts.cpu())
    assert torch.equal(dets.cpu(), out.dets.cpu())
    assert torch.equal(stats.cpu()[:, 0], out.assign_stats.cpu()[:, 0]) and torch.equal(stats.cpu()[:, 3], out.counts.cpu())
    assert px.error() == 0
    px.close()


# ----------------------------------------------------------------------------------------------
# candidate selection: one cluster per image (fused_select_kernel) == the streaming two-pass path
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,batch", [("ssd300_voc_b8", 5), ("ssd_mb2_coco_b64", 7), ("tiny_voc_b3", 3),
                                        ("tiny_sigmoid_b2", 2), ("ssd300_voc_8108_b8", 33)])
def test_fused_cluster_selection_equals_streaming_path(dev, name, batch):
    """The two routes to the candidate lists prune with different (both conservative) gates; everything after --
    exact scores, threshold, top-k, NMS, final top-k -- must come out bit-identical, for every cluster size that
    fits, and so must the row statistics and the sampler's criterion keys."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.pipeline import AnchorPipeline
    w = wl.WORKLOADS[name]
    anchors, gt, scores, locs = wl.make_inputs(w, seed=9, batch=batch)
    lib = N.lib()

    def run(mode):
        N.check(lib.ssd_b200_set_fused_select(mode))
        try:
            pipe = AnchorPipeline(w.cfg())
            target, mask, dets = pipe.step(gt, anchors, scores, locs)
            torch.cuda.synchronize()
            return target.cpu(), mask.cpu(), [d.cpu().clone() for d in dets], pipe.postprocessor.last_anchors.cpu().clone()
        finally:
            N.check(lib.ssd_b200_set_fused_select(-1))

    ref = run(0)
    for mode in (-1, 1, 2, 4, 8):
        got = run(mode)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), (name, mode)
        assert len(got[2]) == len(ref[2])
        for i, (g, r) in enumerate(zip(got[2], ref[2])):
            assert torch.equal(g, r), (name, mode, i)
            assert torch.equal(got[3][i, :g.shape[0]], ref[3][i, :r.shape[0]]), (name, mode, i)


def test_fused_cluster_selection_ties_and_overflow(dev):
    """Heavily tied scores overflow the candidate lists on either route; the exact fallback must give the same rows."""
    from single_shot_detection_b200 import _native as N
    from single_shot_detection_b200.ops import OPS
    lib = N.lib()
    gen = torch.Generator().manual_seed(4)
    b, a, c = 3, 3000, 6
    logits = torch.randint(-2, 3, (b, a, c), generator=gen).float()          # five distinct values: ties en masse
    centre = torch.rand((b, a, 2), generator=gen) * 200
    size = torch.rand((b, a, 2), generator=gen) * 30 + 2
    corners = torch.cat([centre - size / 2, centre + size / 2], dim=-1)
    outs = []
    for mode in (0, -1):
        N.check(lib.ssd_b200_set_fused_select(mode))
        try:
            outs.append([t.cpu() for t in OPS.postprocess(logits.view(b, -1).to(dev), corners.view(b, -1).to(dev), None,
                                                          N.CONVERT_SOFTMAX, 1, N.BOXES_CORNERS, 1.0, 1.0, 0.01, 100, 0.45, 200)])
        finally:
            N.check(lib.ssd_b200_set_fused_select(-1))
    for i in range(b):
        n = int(outs[0][1][i])
        assert n == int(outs[1][1][i])
        assert torch.equal(outs[0][0][i, :n], outs[1][0][i, :n]) and torch.equal(outs[0][2][i, :n], outs[1][2][i, :n])
