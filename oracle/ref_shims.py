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

    class _Attn(torch.nn.Module):
        def __init__(self, dim, num_heads, qkv_bias):
            super().__init__()
            self.num_heads = num_heads
            self.qkv = torch.nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = torch.nn.Linear(dim, dim)

        def forward(self, x):
            b, n, c = x.shape
            qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
            o = torch.nn.functional.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
            return self.proj(o.transpose(1, 2).reshape(b, n, c))

    class _LS(torch.nn.Module):
        def __init__(self, dim, init):
            super().__init__()
            self.gamma = torch.nn.Parameter(init * torch.ones(dim))

        def forward(self, x):
            return x * self.gamma

    class _Mlp(torch.nn.Module):
        def __init__(self, dim, hidden):
            super().__init__()
            self.fc1, self.fc2 = torch.nn.Linear(dim, hidden), torch.nn.Linear(hidden, dim)

        def forward(self, x):
            return self.fc2(torch.nn.functional.gelu(self.fc1(x)))

    class Block(torch.nn.Module):
        """restatement of timm 1.0.24 vision_transformer.Block (pre-LN, LayerScale); drop rates are identity here"""

        def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, proj_drop=0.0, attn_drop=0.0, init_values=None,
                     drop_path=0.0, norm_layer=torch.nn.LayerNorm, **kw):
            super().__init__()
            self.norm1 = norm_layer(dim)
            self.attn = _Attn(dim, num_heads, qkv_bias)
            self.ls1 = _LS(dim, init_values) if init_values else torch.nn.Identity()
            self.norm2 = norm_layer(dim)
            self.mlp = _Mlp(dim, int(dim * mlp_ratio))
            self.ls2 = _LS(dim, init_values) if init_values else torch.nn.Identity()

        def forward(self, x):
            x = x + self.ls1(self.attn(self.norm1(x)))
            return x + self.ls2(self.mlp(self.norm2(x)))

    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")
    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    timm.layers = layers
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.Block = Block
    models.vision_transformer = vt
    timm.models = models
    sys.modules["timm"], sys.modules["timm.layers"] = timm, layers
    sys.modules["timm.models"], sys.modules["timm.models.vision_transformer"] = models, vt


def reference_segformer(name: str, in_channels: int, num_classes: int, dynamic: bool = False):
    """The reference's SegFormerSegmentationModel (weights=None); dynamic=True -> its DynamicMixTransformer encoder."""
    install_timm_shim()
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from geo_deep_learning.models.segmentation.segformer import SegFormerSegmentationModel
    return SegFormerSegmentationModel(encoder=name, in_channels=in_channels, weights=None, num_classes=num_classes,
                                      use_dynamic_encoder=dynamic)


def reference_dofa(img_size: int, embed_dim: int = 768, depth: int = 12, heads: int = 12, out_indices=(4, 6, 10, 11),
                   convert_patch_to_16: bool = False):
    """The reference's DOFAv2 encoder (pretrained=False) on top of the timm Block restatement above."""
    install_timm_shim()
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from geo_deep_learning.models.encoders.dofa_v2 import DOFAv2
    return DOFAv2(img_size=img_size, patch_size=14, embed_dim=embed_dim, depth=depth, num_heads=heads,
                  out_indices=list(out_indices), pretrained=False, drop_path_rate=0.0,
                  convert_patch_to_16=convert_patch_to_16)
