"""B200-native pixel-transform stage of fanlin-rs (the work between decode and
encode: grayscale / inverse, Lanczos3 or Nearest resize, crop, letterbox fill,
Gaussian blur) behind a C ABI (include/fanlin_device.h).

This package holds csrc/ (CUDA kernels + the C-ABI library) and a thin Python
host mirror of the reference's interface for this path, used by tests and
bench.py.  There is no CPU implementation in here: without the built library
and a CUDA device every compute call raises.
"""
from .device import (  # noqa: F401
    Device, DeviceBatch, FanlinError, Job, Plan, lib, lib_path, plan_job, shard_range, FILTER_LANCZOS3, FILTER_NEAREST,
    GRAYSCALE, INVERSE, HAS_DIMS, CROP, TO_RGBA8,
)
from .query import Query  # noqa: F401
from .stage import process_image, process_images, process_gif_frames, make_job  # noqa: F401
