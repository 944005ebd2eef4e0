"""ORACLE (test infrastructure). Restatement of geo_deep_learning/utils/tensors.py:10-35.

Pinned by the reference's own known-answer tests (tests/test_utils_tensors.py:14-50), restated
in tests/test_oracle_cpu.py, and by golden vectors produced by importing the reference module
in the build container (oracle/make_golden.py -> tests/golden/tensors_golden.pt).
"""
from __future__ import annotations

import torch


def normalization(x: torch.Tensor, image_min: float = 0, image_max: float = 255, norm_min: float = 0.0,
                  norm_max: float = 1.0) -> torch.Tensor:
    # utils/tensors.py:18-21: (norm_max - norm_min) * (x - image_min) / (image_max - image_min) + norm_min
    return (norm_max - norm_min) * (x - image_min) / (image_max - image_min) + norm_min


def standardization(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    # utils/tensors.py:31-35: per-channel (x - mean) / std over a (B, C, ...) tensor
    shape = x.shape
    b, c = shape[:2]
    y = (x.reshape(b, c, -1) - mean) / std
    return y.reshape(shape)


def patch_normalise(raw_chw: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """The per-sample step of datasets/wds_dataset.py:230-236: float() -> /255 -> (x-mean)/std with
    (C,1,1)-shaped statistics."""
    x = normalization(raw_chw.float())
    return (x - mean.reshape(-1, 1, 1)) / std.reshape(-1, 1, 1)
