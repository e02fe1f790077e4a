"""ctypes binding of libssd_b200.so (the C ABI declared in include/ssd_b200.h).

There is no fallback: if the shared library is missing or the device is not an sm_100 part the
first call raises.  ``build()`` compiles the library in-tree with nvcc (cross-compiles without a
GPU); it is what ``__graft_entry__.build()`` runs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssd_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ssd_b200.h")

SSD_OK = 0
STATUS_NAMES = {0: "SSD_OK", 1: "SSD_ERR_INVALID_ARGUMENT", 2: "SSD_ERR_MISALIGNED", 3: "SSD_ERR_WORKSPACE",
                4: "SSD_ERR_UNSUPPORTED", 5: "SSD_ERR_CUDA", 6: "SSD_ERR_NO_DEVICE"}

# ssd_box_op
BOX_TO_CORNERS, BOX_TO_CENTROIDS, BOX_TO_CENTROIDS_INPLACE, BOX_ENCODE, BOX_ENCODE_INPLACE, BOX_DECODE, \
    BOX_DECODE_INPLACE, BOX_CENTROIDS_ENCODE_INPLACE, BOX_DECODE_TO_CORNERS = range(9)
# ssd_converter / ssd_box_input
CONVERT_SOFTMAX, CONVERT_SIGMOID, CONVERT_IDENTITY = 0, 1, 2
LOSS_SOFTMAX_CE, LOSS_SIGMOID_FOCAL = 0, 1
BOXES_ENCODED, BOXES_CORNERS = 0, 1


class PostprocessParams(Structure):
    _fields_ = [("batch", c_int32), ("num_anchors", c_int32), ("num_cols", c_int32), ("converter", c_int32),
                ("first_fg_col", c_int32), ("box_input", c_int32), ("xy_scale", c_float), ("wh_scale", c_float),
                ("score_threshold", c_float), ("max_per_class", c_int32), ("overlap_threshold", c_double),
                ("max_total", c_int32), ("det_capacity", c_int32), ("soft_nms", c_int32), ("soft_sigma", c_float),
                ("soft_threshold", c_float), ("resume_after_pass1", c_int32)]


class NativeError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


MAX_ANCHOR_LEVELS = 8
MAX_BOXES_PER_CELL = 16


class AnchorLevel(Structure):
    """SsdAnchorLevel of include/ssd_b200.h."""
    _fields_ = [("cells_x", c_int32), ("cells_y", c_int32), ("x_start", c_float), ("x_end", c_float),
                ("y_start", c_float), ("y_end", c_float), ("num_boxes", c_int32),
                ("wh", c_float * (2 * MAX_BOXES_PER_CELL))]


MAX_PER_CLASS = 512            # kMaxPerClass of csrc/postprocess.cu

_SIGNATURES = {
    "ssd_b200_abi_version": (c_int, []),
    "ssd_b200_last_error": (c_char_p, []),
    "ssd_b200_device_check": (c_int, []),
    "ssd_b200_timing_enable": (None, [c_int]),
    "ssd_b200_timing_report": (c_size_t, [ctypes.c_char_p, c_size_t]),
    "ssd_b200_trace_enable": (c_int, [c_void_p]),
    "ssd_b200_trace_slots": (c_int, []),
    "ssd_pairwise_iou": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ssd_match_per_prediction": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p]),
    "ssd_assign_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ssd_assign_targets": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssd_assign_targets_encoded": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_float,
                                           c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_size_t, c_void_p]),
    "ssd_box_transform": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_float,
                                  c_float, c_float, c_void_p]),
    "ssd_generalized_iou": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ssd_positive_mask": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ssd_b200_launch_count": (ctypes.c_ulonglong, []),
    "ssd_b200_set_stream_ctas_per_sm": (c_int, [c_int]),
    "ssd_b200_set_nms_threads": (c_int, [c_int]),
    "ssd_b200_set_fused_select": (c_int, [c_int]),
    "ssd_mining_keys": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssd_hard_negative_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ssd_hard_negative_mask": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_double,
                                       c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssd_hard_negative_mask_from_keys": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_int, c_double,
                                                 c_void_p, c_void_p, c_void_p]),
    "ssd_postprocess_pass1": (c_int, [POINTER(PostprocessParams), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssd_multibox_loss_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ssd_multibox_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                  c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssd_multibox_loss_giou": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                       c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    "ssd_map_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "ssd_map_append": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "ssd_mean_average_precision": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float,
                                           c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_void_p]),
    "ssd_pack_shard": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ssd_exchange_enable_peer": (c_int, [c_int, c_int]),
    "ssd_exchange_arena_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ssd_exchange_slot_offset": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ssd_pack_exchange": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_void_p), c_int,
                                  c_int, c_int, c_void_p, c_void_p]),
    "ssd_exchange_wait": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "ssd_exchange_open": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p]),
    "ssd_exchange_arena_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "ssd_exchange_arena_free": (c_int, [c_void_p]),
    "ssd_exchange_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "ssd_exchange_peer_close": (c_int, [c_void_p]),
    "ssd_shard_row_words": (c_int, [c_int]),
    "ssd_generate_anchors": (c_int, [POINTER(AnchorLevel), c_int, c_void_p, c_int64, c_void_p]),
    "ssd_postprocess_workspace_bytes": (c_size_t, [POINTER(PostprocessParams)]),
    "ssd_postprocess": (c_int, [POINTER(PostprocessParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssd_nms_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ssd_nms_large_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ssd_nms_large": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_float, c_float, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "ssd_soft_nms": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "ssd_nms": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None
_lock = threading.Lock()


def exported_symbols():
    """Names the header declares (and the library must export)."""
    return sorted(_SIGNATURES)


def build(verbose: bool = False) -> str:
    """Compile libssd_b200.so for sm_100a (nvcc, in-tree)."""
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("building libssd_b200.so failed")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library (loads on first use).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `make -C single_shot_detection_b200/csrc`).  There is no CPU fallback.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)       # AttributeError if the symbol is not exported
                    fn.restype = res
                    fn.argtypes = args
                if handle.ssd_b200_abi_version() != 3:
                    raise ImportError("libssd_b200.so ABI version mismatch")
                _lib = handle
    return _lib


def check(status: int) -> None:
    if status != SSD_OK:
        msg = lib().ssd_b200_last_error()
        raise NativeError(status, msg.decode() if msg else "")


_device_ok = False


def require_device() -> None:
    """Fail loudly unless the current CUDA device can run the sm_100a image."""
    global _device_ok
    if not _device_ok:
        check(lib().ssd_b200_device_check())
        _device_ok = True
