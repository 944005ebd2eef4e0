"""ORACLE support (build container only): import the reference's own SegFormer modules from
/root/reference by satisfying its three `timm.layers` imports (mix_transformer.py:10) with
equivalents — DropPath (per-sample keep mask scaled by 1/keep in train mode, identity in eval),
`to_2tuple`, `trunc_normal_` (torch.nn.init.trunc_normal_, a=-2, b=2).  Nothing here is used on the
GPU box (the reference tree does not exist there)."""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference")


def available() -> bool:
    return (REF / "geo_deep_learning" / "models" / "encoders" / "mix_transformer.py").exists()


def install_timm_shim() -> None:
    if "timm" in sys.modules:
        return

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob: float = 0.0) -> None:
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(t, mean, std, a, b)

    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")
    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    timm.layers = layers
    sys.modules["timm"], sys.modules["timm.layers"] = timm, layers


def reference_segformer(name: str, in_channels: int, num_classes: int):
    """The reference's SegFormerSegmentationModel (weights=None)."""
    install_timm_shim()
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from geo_deep_learning.models.segmentation.segformer import SegFormerSegmentationModel
    return SegFormerSegmentationModel(encoder=name, in_channels=in_channels, weights=None, num_classes=num_classes)
