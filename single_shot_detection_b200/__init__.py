"""B200-native SSD anchor pipeline: a drop-in for the per-image anchor path of
georgymironov/single_shot_detection (target assignment, box coding, hard-negative mining,
post-processor) backed by hand-written sm_100a CUDA kernels behind a C ABI
(include/ssd_b200.h -> libssd_b200.so).  See DESIGN.md.

Importing the package does not load the native library; the first op call does, and raises if
the library is missing or the device is not an sm_100 part.  There is no CPU fallback.
"""

__all__ = ["target_assigner", "matcher", "box_coder", "sampler", "postprocessor", "box_utils", "pipeline",
           "sharding", "workloads", "anchors"]
