"""The timed region of SURVEY.md §8(d) wired through the reference-shaped classes.

``AnchorPipeline.step`` is what ``detection/init.py:108-135`` (``step_fn``) does around the
model in an eval step, minus the model and the loss arithmetic:

    target = target_assigner.encode_ground_truth(ground_truth, priors)          init.py:114
    mask   = sampler(scores, target[..., 4].long())                             multibox_loss.py:49,58
    to_centroids(target[..., :4], inplace=True); encode_box(..., inplace=True)  multibox_loss.py:81-82
    dets   = postprocessor.postprocess((scores, locs), priors)                  init.py:121-122

``step`` takes what the reference's callers hand over (a Python list of ground-truth tensors, CPU
anchors, prediction tensors) and returns the reference's types.  ``step_device`` is the same
sequence on device-resident inputs with no host synchronisation, and ``capture`` records it into
a CUDA graph (launch-bound configs: ~10 kernels of a few microseconds each).
"""
from __future__ import annotations

import functools
import os
from typing import Dict, NamedTuple, Optional

import torch

from . import _devcache, box_utils, sampler as _sampler
from .box_coder import BoxCoder
from .postprocessor import Postprocessor
from .target_assigner import CLASS_INDEX, LOC_INDEX_END, LOC_INDEX_START, PackedGroundTruth, TargetAssigner, pack_ground_truth


class StepOutput(NamedTuple):
    target: torch.Tensor          # [B, A, 6], box columns encoded in place
    mask: torch.Tensor            # [B, A] bool sampler output
    dets: torch.Tensor            # [B, T, 6] padded detections
    counts: torch.Tensor          # [B] int32 rows used in dets
    det_anchors: torch.Tensor     # [B, T] int32 anchor of every detection row
    status: torch.Tensor          # [4] int32 (see ssd_postprocess)
    stats: Optional[torch.Tensor]  # [B, 4] int32 positives, hard negatives, ignored, detections (with shard)
    shard: Optional[torch.Tensor]  # packed exchange buffer, when requested
    assign_stats: Optional[torch.Tensor] = None   # [B, 4] int32 positives, ignored, NaN boxes, G (ssd_assign_targets)
    mining_stats: Optional[torch.Tensor] = None   # [B, 4] int32 positives, negatives, selected, ties (ssd_hard_negative_mask)
    gathered: Optional[torch.Tensor] = None       # [world * capacity, T*6+5] all ranks' shards (gather=True)


class AnchorPipeline:
    def __init__(self, cfg: Dict, workspace_slot: int = 0):
        """``workspace_slot``: pipelines whose steps may run concurrently (two step graphs in flight on two
        streams) need different slots -- the kernels' scratch buffers are cached per slot (ops.workspace_slot).
        ``cfg`` carries the reference's config dict entries (samples/*.py): matched_threshold,
        unmatched_threshold, sampler, ratio, min_neg, xy_scale, wh_scale, eps, score_threshold,
        overlap_threshold, max_per_class, max_total, converter."""
        self.cfg = dict(cfg)
        self.workspace_slot = int(workspace_slot)
        self.target_assigner = TargetAssigner(cfg["matched_threshold"], cfg["unmatched_threshold"])
        self.box_coder = BoxCoder(cfg["xy_scale"], cfg["wh_scale"], cfg.get("eps", 1e-8))
        fn = getattr(_sampler, cfg["sampler"])               # detection/init.py:90-92
        kwargs = {k: v for k, v in {"negative_per_positive_ratio": cfg.get("ratio"),
                                    "min_negative_per_image": cfg.get("min_neg")}.items()
                  if k in fn.__code__.co_varnames}
        self.sampler = functools.partial(fn, **kwargs)
        self.postprocessor = Postprocessor(
            self.box_coder, cfg["score_threshold"],
            {"max_per_class": cfg["max_per_class"], "overlap_threshold": cfg["overlap_threshold"]},
            score_converter=cfg["converter"], max_total=cfg["max_total"])
        self.fuse_encode = False       # True: one pass for to_centroids+encode (same rounding)
        # the assignment writes the box columns already coded for the loss (same bits as the two box passes of
        # multibox_loss.py:81-82, which then disappear from the step); SSD_FUSE_ASSIGN_ENCODE=0 switches it off
        self.fuse_assign_encode = os.environ.get("SSD_FUSE_ASSIGN_ENCODE", "1") != "0"
        # eval step: mining criterion out of the post-processor's first pass (SSD_SHARE_PASS=0: separate kernels)
        self.share_logit_pass = os.environ.get("SSD_SHARE_PASS", "1") != "0"
        # opt-in: the assignment branch starts behind the post-processor's first pass (see _step_device; a single
        # replayed graph gets shorter, 66 -> 59 us on the device timeline, but back-to-back replays measured slower)
        self.assign_after_pass1 = os.environ.get("SSD_ASSIGN_AFTER_PASS1", "0") != "0"
        self.pass1_first = os.environ.get("SSD_PASS1_FIRST", "0") != "0"
        self._graph = None
        self._side = None
        self._copy = None
        self._stream_slots = None
        self._px = None                # sharding.PeerExchange of stream(gather_batch=...)
        self.last_host_stats = None    # [B, 5] int32 host: count, positives, hard negatives, ignored, detections

    # -- the reference-facing call -------------------------------------------------------------
    def step(self, ground_truth, anchors, scores, locs):
        """list of GT tensors, anchors [A,4], scores [B,A*C], locs [B,A*4] (host or device) ->
        (target [B,A,6] with encoded boxes, sampled mask [B,A] bool, list of [n_i,6] detections)."""
        device = torch.device("cuda", torch.cuda.current_device())
        scores = scores if scores.is_cuda else scores.to(device, non_blocking=True)
        locs = locs if locs.is_cuda else locs.to(device, non_blocking=True)
        target = self.target_assigner.encode_ground_truth(ground_truth, anchors)
        if self._shares_logit_pass():
            flight = self.postprocessor.begin_padded((scores, locs), anchors, want_loss_keys=True)
            mask = _sampler.hard_negative_mining_from_keys(flight.loss_keys, target, self.cfg.get("ratio"),
                                                           self.cfg.get("min_neg"), self.target_assigner.last_match)
            self._encode_target_boxes(target, anchors)
            dets = self.postprocessor.to_list(*flight.finish())
        else:
            mask = self._sample_and_encode(target, anchors, scores)
            dets = self.postprocessor.postprocess((scores, locs), anchors)
        return target, mask, dets

    def _shares_logit_pass(self) -> bool:
        # eval step: sampler and post-processor see the same logits, so the post-processor's first pass
        # also emits the mining criterion and the sampler only runs its selection (one read fewer)
        return (self.share_logit_pass and self.cfg["sampler"] == "hard_negative_mining"
                and self.cfg["converter"] == "SOFTMAX")

    def _sample_and_encode(self, target, anchors, scores):
        batch, num_anchors = target.shape[:2]
        classes = target[..., CLASS_INDEX].long()
        mask = self.sampler(scores.view(batch, num_anchors, -1), classes)
        self._encode_target_boxes(target, anchors)
        return mask

    def _encode_target_boxes(self, target, anchors):
        target_locs = target[..., LOC_INDEX_START:LOC_INDEX_END]
        if self.fuse_encode:
            self.box_coder.encode_corners_(target_locs, anchors)
        else:
            box_utils.to_centroids(target_locs, inplace=True)                  # multibox_loss.py:81
            self.box_coder.encode_box(target_locs, anchors, inplace=True)      # multibox_loss.py:82

    # -- the same call over a sequence of host batches, software-pipelined ------------------------
    def stream(self, batches, anchors, depth: int = 2, gather_batch: Optional[int] = None):
        """Generator form of :meth:`step` for a sequence of ``(ground_truth, scores, locs)`` host batches
        (what a data loader with pinned memory hands over): yields ``(target, mask, dets)`` per batch,
        in order.  The host->device copies of batch ``i+1`` run on a copy stream while batch ``i``
        computes, and the padded detections / counts / statistics of every batch are read back into
        pinned host buffers, so ``dets`` is a list of HOST views ``[n_i, 6]``.  ``target`` and ``mask`` stay
        on the device.  Every in-flight batch owns one of ``depth + 1`` static buffer sets with its own
        captured step graph, so what a batch yields stays valid until ``depth`` further batches have
        been yielded.

        ``gather_batch``: under ``torch.distributed`` (one process per GPU, batch sharded by image)
        the global batch size; the detections of all ranks are then exchanged with the one
        all-gather of ``sharding`` before the read-back and ``dets`` covers the global batch."""
        import collections
        from . import sharding
        device = torch.device("cuda", torch.cuda.current_device())
        compute = torch.cuda.current_stream()
        if self._copy is None:
            self._copy = torch.cuda.Stream()
        copy = self._copy
        anchors_dev = _devcache.device_copy(anchors, device)
        if self._stream_slots is None or len(self._stream_slots) != depth + 1:
            self._stream_slots = []
        slots = self._stream_slots  # depth + 1 static buffer sets, one captured step graph each (kept between calls)
        queue = collections.deque()
        issued = 0

        def build_slots(ground_truth, scores, locs):
            # static device inputs + one CUDA graph of the whole step per slot; ground-truth capacity
            # per image is the next power of two >= 64 of what this batch needs
            batch = len(ground_truth)
            need = max([int(g.shape[0]) for g in ground_truth] + [1])
            cap = 64
            while cap < need:
                cap *= 2
            cols = min([int(g.shape[1]) for g in ground_truth if g.dim() == 2 and g.shape[0]] or [6])
            compute.synchronize()
            del slots[:]
            # under torch.distributed the exchange is the peer-memory one (sharding.PeerExchange: the last kernel of
            # every slot's step graph writes the packed shard into every rank's gathered buffer -- no collective
            # call and no allocation per step); it is rebuilt with the slots (collective: every rank sees the same
            # sequence of batch shapes)
            if self._px is not None:
                self._px.close()
                self._px = None
            px = None
            if gather_batch is not None:
                from .ops import det_capacity
                num_cols = scores.numel() // (batch * anchors_dev.shape[0])
                num_fg = num_cols - (1 if self.cfg["converter"] == "SOFTMAX" else 0)
                rows_per_image = det_capacity(num_fg, int(self.cfg["max_per_class"]), int(self.cfg["max_total"] or 0))
                px = self._px = sharding.PeerExchange(gather_batch, rows_per_image, slots=depth + 1)
            for _ in range(depth + 1):
                packed = PackedGroundTruth(torch.zeros((batch * cap, cols), dtype=torch.float32, device=device),
                                           torch.zeros((batch + 1,), dtype=torch.int32, device=device), cap, batch)
                # (the capture warm-up runs on this batch's data: all-zero logits would tie en masse and
                # take the post-processor's slow exact fallback)
                scores_d = scores.to(device=device, dtype=torch.float32).contiguous()
                locs_d = locs.to(device=device, dtype=torch.float32).contiguous()
                pack_ground_truth(ground_truth, device, out=packed)
                runner = AnchorPipeline(self.cfg)
                if px is not None:
                    out = runner.capture(packed, anchors_dev, scores_d, locs_d, exchange=(px, len(slots)))
                else:
                    out = runner.capture(packed, anchors_dev, scores_d, locs_d, shard_capacity=batch)
                rows = gather_batch if gather_batch is not None else batch
                host = torch.empty((rows, out.shard.shape[1]), dtype=torch.float32).pin_memory()
                slots.append({"packed": packed, "scores": scores_d, "locs": locs_d, "runner": runner, "out": out,
                              "host": host, "ready": torch.cuda.Event(), "done": torch.cuda.Event(), "index": len(slots),
                              "key": (tuple(scores.shape), tuple(locs.shape), batch, cols, gather_batch)})

        def batch_key(ground_truth, scores, locs):
            cols = min([int(g.shape[1]) for g in ground_truth if g.dim() == 2 and g.shape[0]] or [6])
            return (tuple(scores.shape), tuple(locs.shape), len(ground_truth), cols, gather_batch)

        def fits(ground_truth, scores, locs):
            return (bool(slots) and slots[0]["key"] == batch_key(ground_truth, scores, locs)
                    and max([int(g.shape[0]) for g in ground_truth] + [0]) <= slots[0]["packed"].max_gt)

        def issue(batch):
            nonlocal issued
            ground_truth, scores, locs = batch
            slot = slots[issued % (depth + 1)]        # its previous use was yielded >= 1 batch ago
            issued += 1
            with torch.cuda.stream(copy):
                pack_ground_truth(ground_truth, device, out=slot["packed"])
                slot["scores"].copy_(scores, non_blocking=True)
                slot["locs"].copy_(locs, non_blocking=True)
                slot["ready"].record(copy)
            compute.wait_event(slot["ready"])
            slot["runner"].replay()
            shard = slot["out"].shard
            if gather_batch is not None:
                px = self._px
                px.wait(slot["index"])                 # every rank's rows of this step have landed in the own arena
                if shard.shape[0] != gather_batch:     # uneven shards: drop the padding rows
                    dets, counts, stats = sharding.unpack_gathered(shard, gather_batch, px.world, slot["out"].dets.shape[1])
                    shard = sharding.pack_shard(dets, counts, stats, gather_batch)
            slot["host"].copy_(shard, non_blocking=True)
            slot["done"].record(compute)
            return slot

        def finish(slot):
            slot["done"].synchronize()
            host = slot["host"]
            t = slot["out"].dets.shape[1]
            ints = host.view(torch.int32)
            self.last_host_stats = ints[:, t * 6: t * 6 + 5]
            dets = host[:, : t * 6].view(host.shape[0], t, 6)
            return slot["out"].target, slot["out"].mask, [dets[i, :n] for i, n in enumerate(ints[:, t * 6].tolist())]

        for batch in batches:
            if not fits(*batch):                      # first batch, new shapes or more boxes per image:
                while queue:                          # drain, then (re)build the static buffers and graphs
                    yield finish(queue.popleft())
                build_slots(*batch)
            queue.append(issue(batch))
            if len(queue) >= depth:
                yield finish(queue.popleft())
        while queue:
            yield finish(queue.popleft())

    def close(self) -> None:
        """Release what :meth:`stream` set up for the multi-GPU exchange (collective under torch.distributed)."""
        if self._px is not None:
            torch.cuda.synchronize()
            self._stream_slots = None
            self._px.close()
            self._px = None

    # -- device-resident, sync-free -------------------------------------------------------------
    def step_device(self, packed: PackedGroundTruth, anchors_dev, scores_dev, locs_dev,
                    shard_capacity: Optional[int] = None, gather: bool = False, exchange=None) -> "StepOutput":
        """``exchange`` = (sharding.PeerExchange, slot): pack the shard and write it into every rank's gathered
        buffer over NVLink peer memory as the last kernel of the step (no NCCL call; capturable)."""
        from . import ops
        with ops.workspace_slot(self.workspace_slot):
            return self._step_device(packed, anchors_dev, scores_dev, locs_dev, shard_capacity, gather, exchange)

    def _step_device(self, packed: PackedGroundTruth, anchors_dev, scores_dev, locs_dev,
                     shard_capacity: Optional[int] = None, gather: bool = False, exchange=None) -> "StepOutput":
        """``shard_capacity``: also pack (dets, counts, stats) into the one-buffer layout
        ``sharding.all_gather_detections`` exchanges (so the packing is part of the graph).
        ``gather``: also run that exchange -- one NCCL all-gather over all ranks -- as the last
        operation of the step (capturable: the collective becomes a node of the step graph).

        The train-side chain (assign -> sampler -> encode) and the post-processor are independent,
        so they are enqueued on two streams (fork / join with events; capturable into one graph
        with two parallel branches)."""
        main = torch.cuda.current_stream()
        side, side2 = self._side_streams()
        side.wait_stream(main)
        batch = packed.batch
        share = self._shares_logit_pass()
        # Branch layout (measured with tools/graph_timeline.py): a kernel behind a CROSS-stream edge of the
        # captured graph started ~10 us after its dependency had finished, a same-stream successor starts at
        # once.  So the train-side chain stays on ONE side stream -- assign -> to_centroids -> encode_box ->
        # selection -- and only the selection has a second (event) dependency, on the post-processor's pass 1.
        coder = self.box_coder if self.fuse_assign_encode else None
        # The assignment shares the SMs with pass 1 when both start together and stretches it from ~11 to ~17 us
        # (tools/graph_timeline.py); `assign_after_pass1` starts its branch behind pass 1 instead (off by default).
        late_assign = share and self.assign_after_pass1
        flight = keyed = None
        # `pass1_first`: pass 1 is CAPTURED before the assignment branch (no dependency between them): the graph then
        # launches it first and its CTAs are resident before the assignment's arrive (tools/graph_timeline.py: the
        # chain pass 1 -> ... -> top-k is ~5 us shorter than when the assignment wins the race for the SMs)
        pass1_first = share and not late_assign and self.pass1_first
        if share and (late_assign or pass1_first):
            flight = self.postprocessor.begin_padded((scores_dev, locs_dev), anchors_dev, want_loss_keys=True)
            keyed = torch.cuda.Event()
            keyed.record(main)
            if late_assign:
                side.wait_event(keyed)
        with torch.cuda.stream(side):
            if exchange is not None:
                exchange[0].open(exchange[1])          # a new launch of the slot: its previous contents are released
            target = self.target_assigner.encode_packed(packed, anchors_dev, box_coder=coder)
            if not share:
                classes = target[..., CLASS_INDEX].long()                      # multibox_loss.py:49 (before the boxes change)
            else:
                classes = None
        if share and not late_assign and not pass1_first:
            flight = self.postprocessor.begin_padded((scores_dev, locs_dev), anchors_dev, want_loss_keys=True)
            keyed = torch.cuda.Event()
            keyed.record(main)
        with torch.cuda.stream(side):
            if coder is None:
                self._encode_target_boxes(target, anchors_dev)                 # touches columns 0-3 only
            if share:
                side.wait_event(keyed)
                mask = _sampler.hard_negative_mining_from_keys(flight.loss_keys, target, self.cfg.get("ratio"),
                                                               self.cfg.get("min_neg"), self.target_assigner.last_match)
            else:
                num_anchors = target.shape[1]
                mask = self.sampler(scores_dev.view(batch, num_anchors, -1), classes)
        if share:
            dets, counts, det_anchors, status = flight.finish()
        else:
            dets, counts, det_anchors, status = self.postprocessor.postprocess_padded((scores_dev, locs_dev), anchors_dev)
        main.wait_stream(side)
        mining = _sampler.hard_negative_mining.last_stats if self.cfg["sampler"] == "hard_negative_mining" else None
        stats, shard = None, None
        if exchange is not None:
            px, slot = exchange
            stats = px.pack_exchange(dets, counts, self.target_assigner.last_stats, mining, slot)
            shard = px.gathered(slot)
        elif shard_capacity is not None:
            from . import sharding
            shard, stats = sharding.pack_shard_device(dets, counts, self.target_assigner.last_stats, mining, shard_capacity)
        gathered = None
        if gather:
            world = torch.distributed.get_world_size()
            gathered = torch.empty((world * shard.shape[0], shard.shape[1]), dtype=shard.dtype, device=shard.device)
            torch.distributed.all_gather_into_tensor(gathered, shard)
        return StepOutput(target, mask, dets, counts, det_anchors, status, stats, shard,
                          self.target_assigner.last_stats, mining, gathered)

    def _side_streams(self):
        if self._side is None:
            self._side = (torch.cuda.Stream(), torch.cuda.Stream())
        return self._side

    def capture(self, packed: PackedGroundTruth, anchors_dev, scores_dev, locs_dev, warmup: int = 2,
                shard_capacity: Optional[int] = None, gather: bool = False, exchange=None,
                concurrent: bool = False) -> "StepOutput":
        """Record ``step_device`` on these (static) buffers into a CUDA graph; returns the outputs
        the replays will keep overwriting.

        ``concurrent``: the graph will be replayed while other step graphs are in flight on other streams: the
        logit-streaming kernels are captured with ONE resident CTA per SM, which leaves shared memory for the
        NMS / selection CTAs of the other steps (slower for a step that runs alone, faster in aggregate)."""
        from . import _native as N
        late = self.assign_after_pass1
        if concurrent:
            N.check(N.lib().ssd_b200_set_stream_ctas_per_sm(int(os.environ.get("SSD_CONCURRENT_CTAS", "1"))))
            self.assign_after_pass1 = False
        try:
            return self._capture(packed, anchors_dev, scores_dev, locs_dev, warmup, shard_capacity, gather, exchange)
        finally:
            self.assign_after_pass1 = late
            if concurrent:
                N.check(N.lib().ssd_b200_set_stream_ctas_per_sm(0))

    def _capture(self, packed, anchors_dev, scores_dev, locs_dev, warmup, shard_capacity, gather, exchange):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                out = self.step_device(packed, anchors_dev, scores_dev, locs_dev, shard_capacity, gather, exchange)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        # The post-processor chain (pass 1 -> pass 2 -> NMS -> top-k) is the longest branch: it is captured
        # on a HIGH-priority stream (the priority becomes an attribute of its kernel nodes), the assignment /
        # sampler / encode branches on default-priority side streams, so that whenever CTAs of both are
        # pending the block scheduler places the critical chain first.
        import os
        prio = os.environ.get("SSD_GRAPH_PRIORITY", "1") != "0"
        with torch.cuda.graph(graph, stream=torch.cuda.Stream(priority=-1) if prio else None):
            out = self.step_device(packed, anchors_dev, scores_dev, locs_dev, shard_capacity, gather, exchange)
        self._graph = graph
        return out

    def replay(self):
        self._graph.replay()


class StepGroup:
    """Several step graphs as the parallel branches of ONE CUDA graph.

    A replay is one ``cudaGraphLaunch`` for ``len(items)`` steps.  As parallel branches (the default) every member
    keeps its own buffers and workspace slot and the members run concurrently: measured, this does NOT raise the
    throughput of steps in flight -- the host needs ~12 us per single-step graph launch, the GPU ~28 us per step
    (profiles/r02_tuning_log.md) -- so ``bench.py --group`` stays at 1.  ``chained=True`` is what the strictly serial leg
    of the bench replays."""

    def __init__(self, items, warmup: int = 2, concurrent: bool = True, chained: bool = False):
        """``items``: [(AnchorPipeline, packed, anchors_dev, scores_dev, locs_dev, kwargs for step_device)]

        ``chained``: the members run strictly ONE AFTER THE OTHER (each step starts when the previous one has finished --
        the graph is a chain of step sub-graphs on one stream), so a replay is what ``len(items)`` replays of single-step
        graphs on one stream are, minus the graph-to-graph launch latency the stream would expose between them."""
        from . import _native as N
        slots = [it[0].workspace_slot for it in items]
        assert chained or len(set(slots)) == len(slots), "members of a step group run concurrently: one workspace slot each"
        concurrent = concurrent and not chained
        if chained:
            # a step that runs alone: 256-thread NMS CTAs finish ~1 us sooner at the headline size (640 segments, about
            # four per SM); with thousands of segments (the COCO configurations) 128 threads pack slightly better --
            # measured after the last GPU run of the round, so the chain keeps the setting that was validated
            N.check(N.lib().ssd_b200_set_nms_threads(256))
        if concurrent:
            N.check(N.lib().ssd_b200_set_stream_ctas_per_sm(int(os.environ.get("SSD_CONCURRENT_CTAS", "1"))))
        late = [it[0].assign_after_pass1 for it in items]
        try:
            if concurrent:
                for it in items:
                    it[0].assign_after_pass1 = False
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    for pipe, packed, anchors_dev, scores_dev, locs_dev, kw in items:
                        pipe.step_device(packed, anchors_dev, scores_dev, locs_dev, **kw)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            branches = [torch.cuda.Stream(priority=-1) for _ in items]
            self.outs = []
            root = torch.cuda.Stream()
            if chained:
                with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(priority=-1)):
                    for pipe, packed, anchors_dev, scores_dev, locs_dev, kw in items:
                        self.outs.append(pipe.step_device(packed, anchors_dev, scores_dev, locs_dev, **kw))
                return
            with torch.cuda.graph(self.graph, stream=root):
                for br, (pipe, packed, anchors_dev, scores_dev, locs_dev, kw) in zip(branches, items):
                    br.wait_stream(root)
                    with torch.cuda.stream(br):
                        self.outs.append(pipe.step_device(packed, anchors_dev, scores_dev, locs_dev, **kw))
                for br in branches:
                    root.wait_stream(br)
        finally:
            if chained:
                N.check(N.lib().ssd_b200_set_nms_threads(0))
            for it, l in zip(items, late):
                it[0].assign_after_pass1 = l
            if concurrent:
                N.check(N.lib().ssd_b200_set_stream_ctas_per_sm(0))

    def replay(self):
        self.graph.replay()


def matched_stats(assign_stats: torch.Tensor, mining_stats: Optional[torch.Tensor], counts: torch.Tensor):
    """[B,4] int32 {positives, hard negatives selected, ignored, detections}."""
    out = torch.zeros((assign_stats.shape[0], 4), dtype=torch.int32, device=assign_stats.device)
    out[:, 0] = assign_stats[:, 0]
    if mining_stats is not None:
        out[:, 1] = mining_stats[:, 2]
    out[:, 2] = assign_stats[:, 1]
    out[:, 3] = counts
    return out
