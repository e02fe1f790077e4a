"""CPU-side checks: the C-ABI library loads and exports every symbol include/ssd_b200.h declares,
the ctypes binding mirrors the header, the product package has no route into oracle/, and the
reference-shaped Python surface keeps the reference's names.  No kernel is launched here."""
import ast
import ctypes
import inspect
import os
import re

import pytest

from single_shot_detection_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "single_shot_detection_b200")


def _header_symbols():
    text = open(N.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"SSD_API\s+[\w\s\*]+?\b(ssd_\w+)\s*\(", text)))


def test_header_declares_the_bound_symbols():
    assert _header_symbols() == N.exported_symbols()


def test_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(N.LIB_PATH), "libssd_b200.so missing: run __graft_entry__.build()"
    handle = ctypes.CDLL(N.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(handle, name), f"{name} is declared in ssd_b200.h but not exported"
    handle.ssd_b200_abi_version.restype = ctypes.c_int
    assert handle.ssd_b200_abi_version() == 3
    # only the declared entry points are visible (the kernels and helpers are hidden)
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(ln.split()[-1] for ln in out.splitlines() if " T " in ln and "ssd_" in ln.split()[-1])
    assert exported == _header_symbols()


def test_postprocess_params_struct_matches_header_layout():
    # int32 x6, float x3, int32, double, int32 x3, float x2, int32 -> the double sits at offset 40 on every LP64 ABI
    assert N.PostprocessParams.overlap_threshold.offset == 40
    assert N.PostprocessParams.soft_nms.offset == 56
    assert ctypes.sizeof(N.PostprocessParams) == 72
    assert ctypes.sizeof(N.AnchorLevel) == 7 * 4 + 2 * N.MAX_BOXES_PER_CELL * 4


def test_argument_validation_without_a_device():
    """Entry points validate their arguments before touching CUDA, so these run without a GPU."""
    lib = N.lib()
    assert lib.ssd_box_transform(99, None, 4, None, 4, None, 1, 1, 1.0, 1.0, 0.0, None) == 1
    assert b"bad op" in lib.ssd_b200_last_error()
    assert lib.ssd_assign_targets(None, None, 6, None, 0, -1, 10, 0.5, 0.5, 1, None, None, None, None, 0, None) == 1
    assert lib.ssd_hard_negative_mask(None, None, None, 1, 8, 3, 3.0, 1, 5.0, None, None, None, 0, None) == 1
    p = N.PostprocessParams()
    p.batch, p.num_anchors, p.num_cols, p.converter = 1, 8, 3, 7
    assert lib.ssd_postprocess_workspace_bytes(ctypes.byref(p)) == 0
    assert b"score_converter" in lib.ssd_b200_last_error()
    # empty work is a no-op, not an error
    assert lib.ssd_positive_mask(None, 0, None, None) == 0


def test_product_package_never_imports_the_oracle():
    for fn in os.listdir(PKG):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(PKG, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names), f"{fn} imports the oracle"


def test_reference_surface_names_and_signatures():
    from single_shot_detection_b200 import box_coder, box_utils, matcher, postprocessor, sampler, target_assigner
    # detection/target_assigner.py:7-14
    assert (target_assigner.LOC_INDEX_START, target_assigner.LOC_INDEX_END, target_assigner.CLASS_INDEX,
            target_assigner.SCORE_INDEX, target_assigner.TARGET_SIZE) == (0, 4, 4, 5, 6)
    assert (target_assigner.NEGATIVE_CLASS, target_assigner.IGNORE_CLASS) == (0, -1)
    assert (matcher.NOT_MATCHED, matcher.IGNORE) == (-2, -1)
    # detection/init.py:90-92 filters the sampler config through __code__.co_varnames
    names = sampler.hard_negative_mining.__code__.co_varnames
    assert "negative_per_positive_ratio" in names and "min_negative_per_image" in names
    assert list(inspect.signature(sampler.naive_sampler).parameters) == ["predictions", "target_classes"]
    assert list(inspect.signature(target_assigner.TargetAssigner.encode_ground_truth).parameters) == \
        ["self", "ground_truth", "anchors"]
    assert list(inspect.signature(box_coder.BoxCoder.__init__).parameters)[:4] == ["self", "xy_scale", "wh_scale", "eps"]
    assert list(inspect.signature(box_coder.BoxCoder.encode_box).parameters) == ["self", "boxes", "priors", "inplace"]
    assert list(inspect.signature(box_coder.BoxCoder.decode_box).parameters) == ["self", "boxes", "priors", "inplace"]
    assert list(inspect.signature(postprocessor.Postprocessor.__init__).parameters) == \
        ["self", "box_coder", "score_threshold", "nms", "score_converter", "max_total"]
    assert list(inspect.signature(box_utils.nms).parameters) == \
        ["boxes", "scores", "overlap_threshold", "score_threshold", "max_per_class", "soft", "sigma"]
    with pytest.raises(ValueError):                       # detection/postprocessor.py:22
        postprocessor.Postprocessor(box_coder.BoxCoder(10, 5), .01, {"max_per_class": 100, "overlap_threshold": .45},
                                    score_converter="TANH")


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from single_shot_detection_b200 import box_utils
    with pytest.raises(TypeError):
        box_utils.to_corners(torch.zeros(3, 4))
    with pytest.raises(TypeError):
        box_utils.iou(torch.zeros(3, 4), torch.zeros(2, 4))


def test_anchor_generator_shapes_match_reference_tables():
    """Host half of the device anchor generator: the per-level (w, h) scalars and the linspace end points
    reproduce the tables the reference's own generators produced (tests/golden/anchors.npz); the cell
    centres themselves are written by the kernel (-m gpu)."""
    import numpy as np
    import torch
    import golden_io as gio
    from single_shot_detection_b200 import workloads as wl
    z = gio.load("anchors.npz")
    for name in z.files:
        w = wl.WORKLOADS[name]
        table = z[name]
        first = 0
        for gen, cells in zip(wl.build_anchor_generators(w), w.fmaps):
            lvl = gen._level((w.img, w.img), (cells, cells))
            n = cells * cells * gen.num_boxes
            block = table[first: first + n].reshape(cells, cells, gen.num_boxes, 4)
            wh = np.array(list(lvl.wh)[: 2 * gen.num_boxes], dtype=np.float32).reshape(-1, 2)
            assert np.array_equal(block[0, 0, :, 2:], wh), name
            assert np.float32(lvl.x_start) == block[0, 0, 0, 0] and np.float32(lvl.x_end) == block[0, -1, 0, 0], name
            assert np.float32(lvl.y_start) == block[0, 0, 0, 1] and np.float32(lvl.y_end) == block[-1, 0, 0, 1], name
            first += n
        assert first == table.shape[0]


@pytest.mark.skipif(not os.path.isdir(os.environ.get("SSD_REFERENCE_ROOT", "/root/reference") + "/samples"),
                    reason="the reference's samples/*.py exist in the build container only")
def test_every_reference_sample_config_constructs_the_mirrored_classes():
    """detection/init.py:90-97 builds sampler / BoxCoder / MultiboxLoss / Postprocessor / TargetAssigner from the
    dicts of a samples/*.py file: every sample the reference ships must construct the mirrored classes as-is
    (construction is host-only: no device needed)."""
    import functools
    import glob
    import runpy
    from single_shot_detection_b200 import sampler as samplers
    from single_shot_detection_b200.box_coder import BoxCoder
    from single_shot_detection_b200.multibox_loss import MultiboxLoss
    from single_shot_detection_b200.postprocessor import Postprocessor
    from single_shot_detection_b200.target_assigner import TargetAssigner
    root = os.environ.get("SSD_REFERENCE_ROOT", "/root/reference")
    files = sorted(glob.glob(os.path.join(root, "samples", "*.py")))
    assert len(files) >= 13
    for path in files:
        cfg = runpy.run_path(path)
        fn = getattr(samplers, cfg["sampler"]["name"])                                     # init.py:90
        kwargs = {k: v for k, v in cfg["sampler"].items() if k in fn.__code__.co_varnames}  # init.py:91
        smp = functools.partial(fn, **kwargs)
        coder = BoxCoder(**cfg["box_coder"])                                               # init.py:94
        crit = MultiboxLoss(sampler=smp, box_coder=coder, **cfg["loss"])                   # init.py:95
        post = Postprocessor(coder, **cfg["postprocess"])                                  # init.py:96
        assigner = TargetAssigner(**cfg["target_assigner"])                                # init.py:97
        assert post.box_coder is coder and post.score_threshold == cfg["postprocess"]["score_threshold"], path
        assert post.max_total == cfg["postprocess"].get("max_total"), path
        assert assigner.matched_threshold == cfg["target_assigner"]["matched_threshold"], path
        assert crit is not None


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver runs it before our arm, on the box's host cores): one JSON line with the
    metric / config of our arm, `impl: reference`, the CPU baseline it timed and an `e2e` block without copies.  A tiny
    sample here (2 images, 1 step): the plumbing, not the number."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "1", "--cpu-sample-images", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "images/sec target-assign+NMS" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and line["vs_baseline"] is None and line["dtype"] == "f32"
    assert line["config"]["workload"] == "ssd300_voc_b32" and line["config"]["anchors"] == 8732
    for key in ("l2", "steps_in_flight", "steps_per_graph_launch", "device_path"):      # the same keys as the GPU arm
        assert key in line["config"], key
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "2 of 32 images" in cb["sample"]
    assert cb["one_thread"]["cores"] == 1 and set(cb["stage_ms_per_sample"]) >= {"encode_ground_truth", "sampler", "postprocess"}
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
