"""gdl_b200 — B200-native (sm_100a) hot path for NRCan/geo-deep-learning segmentation models.

Host-side Python mirrors the reference's model/task plugin surface; all arithmetic runs in
hand-written CUDA kernels reached through the C ABI in include/gdl_b200.h.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
