"""B200-native DOFA-v2 encoder (forward) + the DOFA segmentation model.

`DOFAv2` mirrors geo_deep_learning/models/encoders/dofa_v2.py (constructor keywords, `state_dict` keys
`patch_embed.weight_generator.transformer_encoder.layers.0.self_attn.in_proj_weight`, `blocks.3.attn.qkv.weight`,
`blocks.3.ls1.gamma`, `pos_embed`, `cls_token`, ...; `create_dofa_base/large`), `DOFASegmentationModel` mirrors
models/segmentation/dofa.py:24-107 (encoder + MultiLevelNeck + UperNetDecoder + SegmentationHead + FCNHead,
`forward(x, wavelengths) -> SegmentationOutput(out, aux)`).

Two routes through the encoder:

* FROZEN (the shipped configuration: `freeze_layers: ["encoder"]`, configs/dofa_config_RGB.yaml:57) — forward only,
  `_features_nhwc`.  Every matmul (weight-generator transformer layer, dynamic patch embedding, qkv / proj / MLP of the
  12 ViT blocks, q.k^T and P.V) is the tcgen05 GEMM kernel; LayerNorm / softmax / GELU / LayerScale / residual are the
  fused epilogues and row kernels of libgdlb200.so.  Host-side torch is used only for glue on tiny tensors: the sin/cos
  of <= 12 wavelengths, concatenating the 128+C+1 generator tokens, and re-laying the generated (C, 14*14*D) weights as
  OIHW (x 0.01) before the packing kernel.
* TRAINABLE — `run_train` / `backward` (hand-written backward on the engine, like SegFormer): the linears go through
  `Engine.conv_raw` / `conv_backward` (dgrad + wgrad GEMMs), attention backward is `dP = dO.V^T`, softmax backward,
  `dQ = dS.K`, `dV = P^T dO`, `dK = dS^T q`; GELU and LayerScale (+ timm's DropPath as a per-sample factor) run as their
  own kernels so the pre-activation and the un-scaled branch output survive for the backward.  The weight GENERATOR
  (130 tokens x 128 channels, < 0.01 % of the step's FLOPs) is differentiated by torch autograd in fp32: its output —
  the patch-embedding weights — gets its gradient from the wgrad kernel, and `torch.autograd.grad` carries it to the
  generator's parameters.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..engine import Act, Engine
from .upernet import FCNHead, MultiLevelNeck, SegmentationHead, UperNetDecoder, UperNetSegmentor, _UperNetFn


class SegmentationOutput(NamedTuple):
    out: torch.Tensor
    aux: torch.Tensor | None


def _sincos_1d(embed_dim: int, pos: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    omega = torch.arange(embed_dim // 2, dtype=torch.float32, device=pos.device) / (embed_dim / 2.0)
    omega = (1.0 / 10000 ** omega).to(dtype)
    out = torch.einsum("m,d->md", pos.reshape(-1).to(dtype), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1)


def _sincos_2d(embed_dim: int, grid: int) -> torch.Tensor:
    gh, gw = torch.meshgrid(torch.arange(grid), torch.arange(grid), indexing="ij")
    pe = torch.cat([_sincos_1d(embed_dim // 2, gh), _sincos_1d(embed_dim // 2, gw)], dim=1)
    return torch.cat([torch.zeros(1, embed_dim), pe], dim=0)


class _FCRes(nn.Module):
    def __init__(self, n: int = 128) -> None:
        super().__init__()
        self.w1, self.w2 = nn.Linear(n, n), nn.Linear(n, n)


class _WeightGenerator(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, embed_dim: int) -> None:
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=input_dim, nhead=4, activation="gelu", norm_first=False,
                                           batch_first=False, dropout=0.0)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=1, enable_nested_tensor=False)
        self.fc_weight = nn.Linear(input_dim, output_dim)
        self.fc_bias = nn.Linear(input_dim, embed_dim)
        self.weight_tokens = nn.Parameter(torch.empty(128, input_dim))
        self.bias_token = nn.Parameter(torch.empty(1, input_dim))
        nn.init.normal_(self.weight_tokens, std=0.02)
        nn.init.normal_(self.bias_token, std=0.02)


class _Embedding(nn.Module):
    def __init__(self, kernel_size: int, embed_dim: int) -> None:
        super().__init__()
        self.weight_generator = _WeightGenerator(128, kernel_size * kernel_size * embed_dim, embed_dim)
        self.fclayer = _FCRes(128)
        for m in self.modules():  # DOFAv2Embedding._init_weights
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                m.bias.data.fill_(0.01)


class _LayerScale(nn.Module):
    def __init__(self, dim: int, init: float) -> None:
        super().__init__()
        self.gamma = nn.Parameter(init * torch.ones(dim))


class _Attn(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim: int, ratio: float) -> None:
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, int(dim * ratio)), nn.Linear(int(dim * ratio), dim)


class _ViTBlock(nn.Module):
    def __init__(self, dim: int, ratio: float, init_values: float) -> None:
        super().__init__()
        self.norm1, self.attn, self.ls1 = nn.LayerNorm(dim), _Attn(dim), _LayerScale(dim, init_values)
        self.norm2, self.mlp, self.ls2 = nn.LayerNorm(dim), _Mlp(dim, ratio), _LayerScale(dim, init_values)


class _SavedBlock:
    """tensors one ViT block (or the whole encoder) keeps for the hand-written backward"""

    def __init__(self, **kw) -> None:
        self.__dict__.update(kw)


_PACKED: dict = {}  # id(weight) -> (version, dtype, packed 16-bit operand): frozen encoder weights are packed once


def _packed(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    key = id(w)
    hit = _PACKED.get(key)
    if hit is not None and hit[0] == w._version and hit[1] == dtype and hit[2].device == w.device and hit[3]() is w:
        return hit[2]
    n, k = w.shape
    wp = ops.pack_conv_weight(w.detach().float().contiguous().view(n, k, 1, 1), dtype)
    import weakref
    _PACKED[key] = (w._version, dtype, wp, weakref.ref(w))
    return wp


def _lin(x2d: torch.Tensor, lin, *, out_dtype=None, relu=False, gelu=False, residual=None, oscale=None, w=None, b=None):
    """rows (M, K) 16-bit -> (M, N): the tcgen05 GEMM with bias / activation / LayerScale / residual epilogue"""
    w = lin.weight if w is None else w
    b = (lin.bias if lin is not None else None) if b is None else b
    m, k = x2d.shape
    n = w.shape[0]
    wp = _packed(w, x2d.dtype)
    res4 = residual.view(1, 1, m, n) if residual is not None else None
    y = ops.conv2d_fwd([x2d.view(1, 1, m, k)], wp, n, 1, 1, 0, 0, out_dtype=out_dtype or x2d.dtype,
                       bias=b.detach().float() if b is not None else None, relu=relu, gelu=gelu, residual=res4,
                       oscale=oscale.detach().float() if oscale is not None else None)
    return y.view(m, n)


def _mha(qkv: torch.Tensor, b: int, n: int, heads: int, c: int, dt) -> torch.Tensor:
    """qkv (b*n, 3c) 16-bit with columns [q | k | v] -> softmax(q k^T d^-1/2) v as (b*n, c)"""
    d = c // heads
    if ops.option("mha_flash") and d == 64:
        return ops.mha_flash_fwd(qkv, b, n, heads, d ** -0.5)  # one kernel, no score / probability tensors (csrc/sra_attention.cu)
    # key rows padded to a multiple of 64: the P.V GEMM then contracts over 128-byte (64-key) swizzled rows and the
    # q.k^T GEMM writes whole lp-wide score rows through the TMA-store epilogue.  Score columns >= n hold products with
    # the next image's keys (or TMA zero fill): the softmax kernel ignores them and writes zeros there.
    lp = (n + 63) // 64 * 64
    q4 = qkv.view(b, 1, n, 3 * c)
    scores = torch.empty((b, 1, n, heads * lp), dtype=dt, device=qkv.device)
    # one grouped launch per product: head g reads q / k / v columns shifted by g*d and writes scores / P columns by g*lp
    ops.conv2d_fwd([q4[..., 0:d]], qkv[:, c:c + d], lp, 1, 1, 0, 0, out=scores[..., 0:lp], w_rows_per_img=n,
                   groups=(heads, d, d, lp))
    p = ops.softmax_fwd(scores.view(b, n, heads, lp), d ** -0.5, n)
    p4 = p.view(b, 1, n, heads * lp)
    o = torch.empty((b, 1, n, c), dtype=dt, device=qkv.device)
    ops.conv2d_fwd([p4[..., 0:lp]], qkv[:, 2 * c:2 * c + d], d, 1, 1, 0, 0, out=o[..., 0:d], w_rows_per_img=n,
                   w_mn_major=True, groups=(heads, lp, d, d))
    return o.view(b * n, c)


class DOFAv2(nn.Module):
    def __init__(self, encoder_name: str = "dofa_base", img_size=224, patch_size: int = 14, embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, drop_rate: float = 0.0,
                 drop_path_rate: float = 0.1, out_indices: list[int] | None = None, init_values: float = 1e-5, *,
                 convert_patch_to_16: bool = False, pretrained: bool = False,
                 compute_dtype: torch.dtype = torch.bfloat16) -> None:
        super().__init__()
        if pretrained:
            raise ValueError("pretrained=True downloads from HuggingFace; load the tensors with load_state_dict")
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.encoder_name, self.img_size, self.patch_size = encoder_name, img_size, patch_size
        self.embed_dim, self.depth, self.num_heads = embed_dim, depth, num_heads
        # convert_patch_to_16 (dofa_v2.py:168-176,220): the generated 14x14 kernels are resampled to 16x16 (bicubic,
        # align_corners=False) and applied with stride 16 — the kernels are (D, C, 14, 14), a sub-MB tensor: torch glue
        self.convert_patch_to_16 = convert_patch_to_16
        self.conv_kernel = 16 if convert_patch_to_16 else patch_size
        self.num_patches = (img_size[0] // self.conv_kernel) * (img_size[1] // self.conv_kernel)
        self.out_indices = list(out_indices) if out_indices is not None else [depth - 1]
        self.patch_embed = _Embedding(patch_size, embed_dim)
        self.pos_embed = nn.Parameter(_sincos_2d(embed_dim, int(self.num_patches ** 0.5)).unsqueeze(0), requires_grad=False)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        nn.init.normal_(self.cls_token, std=0.02)
        self.blocks = nn.ModuleList(_ViTBlock(embed_dim, mlp_ratio, init_values) for _ in range(depth))
        self.norm = nn.LayerNorm(embed_dim)
        self.compute_dtype = compute_dtype
        # timm DropPath rates of the blocks (dofa_v2.py:248: linspace(0, drop_path_rate, depth)); only the TRAINING route
        # draws masks.  `drop_path_masks` (list of (B,) keep/keep_prob factors per block, or None) overrides the draw.
        self.drop_path_rates = [float(v) for v in torch.linspace(0, drop_path_rate, depth)]
        self.drop_path_masks: list | None = None
        self._saved = None

    # ---------------------------------------------------------------------------------- weight generator
    @torch.no_grad()
    def _dynamic_weights(self, wavelengths: torch.Tensor, c: int):
        """wavelengths (C,) in um -> (OIHW fp32 weights (D, C, k, k) * 0.01, bias (D,) * 0.01)   (dofa_v2.py:148-166)"""
        dt = self.compute_dtype
        pe, wg = self.patch_embed, self.patch_embed.weight_generator
        waves = _sincos_1d(128, wavelengths.float() * 1000).contiguous()                       # (C,128) fp32
        h1 = _lin(ops.cast_f32(waves, dt), pe.fclayer.w1, relu=True)
        h2 = _lin(h1, pe.fclayer.w2, relu=True, out_dtype=torch.float32)
        waves = ops.add_nhwc(waves.view(1, 1, c, 128), h2.view(1, 1, c, 128)).view(c, 128)     # x + relu(w2 relu(w1 x))
        x = torch.cat([wg.weight_tokens.detach().float(), waves, wg.bias_token.detach().float()], 0).contiguous()
        n = x.shape[0]
        layer = wg.transformer_encoder.layers[0]
        sa = layer.self_attn
        qkv = _lin(ops.cast_f32(x, dt), None, w=sa.in_proj_weight, b=sa.in_proj_bias)          # (n, 384)
        att = _mha(qkv, 1, n, 4, 128, dt)
        y = _lin(att, sa.out_proj, residual=x, out_dtype=torch.float32)                        # x + SA(x)
        x1, _ = ops.layernorm_fwd(y, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, torch.float32, False)
        f = _lin(ops.cast_f32(x1, dt), layer.linear1, gelu=True)
        y = _lin(f, layer.linear2, residual=x1, out_dtype=torch.float32)
        x2, _ = ops.layernorm_fwd(y, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, torch.float32, False)
        wt = 128
        wsrc = ops.add_nhwc(x2[wt:wt + c].contiguous().view(1, 1, c, 128), waves.view(1, 1, c, 128)).view(c, 128)
        weights = _lin(ops.cast_f32(wsrc, dt), wg.fc_weight, out_dtype=torch.float32)          # (C, k*k*D)
        bias = _lin(ops.cast_f32(x2[n - 1:n].contiguous(), dt), wg.fc_bias, out_dtype=torch.float32)
        k = self.patch_size
        w_oihw = (weights.view(c, k, k, self.embed_dim).permute(3, 0, 1, 2) * 0.01).contiguous()
        if self.convert_patch_to_16:
            w_oihw = F.interpolate(w_oihw, size=(16, 16), mode="bicubic", align_corners=False).contiguous()
        return w_oihw, (bias.view(self.embed_dim) * 0.01).contiguous()

    # ---------------------------------------------------------------------------------- forward
    def forward_features(self, x: torch.Tensor, wavelengths: torch.Tensor) -> list[torch.Tensor]:
        """x (B,C,H,W) float, wavelengths (C,) or (B,C) -> NHWC 16-bit maps (B, h, w, D) at out_indices"""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("DOFAv2 is forward-only here: freeze the encoder (freeze_layers=['encoder']) "
                                      "or call it under torch.no_grad()")
        with torch.no_grad():
            return self._features(x, wavelengths)

    def _features(self, x: torch.Tensor, wavelengths: torch.Tensor) -> list[torch.Tensor]:
        if wavelengths.dim() == 2:
            if not torch.allclose(wavelengths, wavelengths[0:1].expand_as(wavelengths)):
                raise ValueError("DOFA cannot handle different wavelengths within a batch")
            wavelengths = wavelengths[0]
        c = x.shape[1]
        img = ops.normalize_to_nhwc(x.contiguous().float(), True, self.compute_dtype, (c + 7) // 8 * 8)
        return self._features_nhwc(img, c, wavelengths)

    def _features_nhwc(self, img: torch.Tensor, c: int, wavelengths: torch.Tensor) -> list[torch.Tensor]:
        """img: NHWC 16-bit tile batch (B, H, W, ld >= c) as produced by the normalise kernel"""
        dt = self.compute_dtype
        b, hh, ww = img.shape[:3]
        d, k = self.embed_dim, self.conv_kernel
        w_oihw, bias = self._dynamic_weights(wavelengths, c)
        kk = k * k * c
        kpad = (kk + 63) // 64 * 64
        col = ops.im2col(img, c, k, k, k, 1, kpad)
        wp = ops.pack_conv_weight(w_oihw, dt, 0, kpad)
        patch = ops.conv2d_fwd([col], wp, d, 1, 1, 0, 0, bias=bias)                           # (B, h, w, D)
        hp, wpx = patch.shape[1:3]
        p_tok = hp * wpx
        if p_tok + 1 != self.pos_embed.shape[1]:
            raise ValueError(f"image {hh}x{ww} gives {p_tok} patches but pos_embed has {self.pos_embed.shape[1] - 1}")
        tokens = ops.vit_assemble_tokens(patch.view(b, p_tok, d), self.pos_embed[0].contiguous(),
                                         self.cls_token.detach().float().view(d).contiguous())
        n = p_tok + 1
        feats = []
        for i, blk in enumerate(self.blocks):
            t2 = tokens.view(b * n, d)
            a, _ = ops.layernorm_fwd(t2, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, dt, False)
            qkv = _lin(a, blk.attn.qkv)
            o = _mha(qkv, b, n, self.num_heads, d, dt)
            t2 = _lin(o, blk.attn.proj, residual=t2, oscale=blk.ls1.gamma, out_dtype=torch.float32)
            a, _ = ops.layernorm_fwd(t2, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, dt, False)
            f = _lin(a, blk.mlp.fc1, gelu=True)
            t2 = _lin(f, blk.mlp.fc2, residual=t2, oscale=blk.ls2.gamma, out_dtype=torch.float32)
            tokens = t2.view(b, n, d)
            if i in self.out_indices:
                feats.append(ops.vit_extract_feature(tokens, dt).view(b, hp, wpx, d))
        # reference quirk kept: the final norm is only applied when depth-1 was requested but not tapped (never)
        return feats

    # ---------------------------------------------------------------------------------- training route
    def trainable(self) -> bool:
        return any(p.requires_grad for p in self.parameters())

    def _generator_params(self) -> list[torch.Tensor]:
        return [p for p in self.patch_embed.parameters() if p.requires_grad]

    def _dynamic_weights_autograd(self, wavelengths: torch.Tensor, c: int):
        """dofa_v2.py:148-166 in torch (parameter dtype) under autograd: (D,C,k,k) weights * 0.01, (D,) bias * 0.01."""
        pe, wg = self.patch_embed, self.patch_embed.weight_generator
        pdt = wg.weight_tokens.dtype
        with torch.enable_grad():
            waves = _sincos_1d(128, wavelengths.to(pdt) * 1000, pdt)  # fp32 parameters: the reference's arithmetic
            y = F.relu(F.linear(waves, pe.fclayer.w1.weight, pe.fclayer.w1.bias))
            waves = waves + F.relu(F.linear(y, pe.fclayer.w2.weight, pe.fclayer.w2.bias))
            x = torch.cat([wg.weight_tokens, waves, wg.bias_token], 0)
            layer = wg.transformer_encoder.layers[0]  # post-norm encoder layer, dropout 0 (written out: no fast path)
            sa = layer.self_attn
            n = x.shape[0]
            q, k_, v = F.linear(x, sa.in_proj_weight, sa.in_proj_bias).view(n, 3, sa.num_heads, -1).unbind(1)
            att = torch.softmax(torch.einsum("nhd,mhd->hnm", q, k_) * q.shape[-1] ** -0.5, -1)
            o = torch.einsum("hnm,mhd->nhd", att, v).reshape(n, -1)
            x = layer.norm1(x + sa.out_proj(o))
            x = layer.norm2(x + layer.linear2(F.gelu(layer.linear1(x))))
            weights = wg.fc_weight(x[128:-1] + waves)
            bias = wg.fc_bias(x[-1])
            kk = self.patch_size
            w_oihw = weights.view(c, kk, kk, self.embed_dim).permute(3, 0, 1, 2) * 0.01
            if self.convert_patch_to_16:
                w_oihw = F.interpolate(w_oihw, size=(16, 16), mode="bicubic", align_corners=False)
            return w_oihw, bias.view(self.embed_dim) * 0.01

    def _drop_path_factors(self, i: int, b: int, dev, training: bool):
        """(B,) fp32 keep-mask / keep-probability of block i (both branches draw their own), or None."""
        if self.drop_path_masks is not None:
            return self.drop_path_masks[i]
        rate = self.drop_path_rates[i]
        if not training or rate <= 0.0:
            return None, None
        keep = 1.0 - rate
        draw = torch.bernoulli(torch.full((2, b), keep, dtype=torch.float32, device=dev)) / keep
        return draw[0].contiguous(), draw[1].contiguous()

    def _lin_train(self, eng: Engine, x2d: torch.Tensor, lin: nn.Linear, out_dtype=None):
        m, k = x2d.shape
        a = Act(x2d.view(1, 1, m, k))
        rc = eng.conv_raw([a], lin.weight, 1, 0, bias=lin.bias, out_dtype=out_dtype,
                          wshape=(lin.weight.shape[0], lin.weight.shape[1], 1, 1))
        return rc, a

    def run_train(self, eng: Engine, img: torch.Tensor, c: int, wavelengths: torch.Tensor) -> list[Act]:
        """img: NHWC 16-bit tiles (B,H,W,ld >= c).  Returns the tapped maps as gradient-carrying activations; call
        `backward(eng)` after their consumers have registered their gradients."""
        if wavelengths.dim() == 2:
            if not torch.allclose(wavelengths, wavelengths[0:1].expand_as(wavelengths)):
                raise ValueError("DOFA cannot handle different wavelengths within a batch")
            wavelengths = wavelengths[0]
        dt, acc = eng.dtype, eng.acc_dtype
        b, hh, ww = img.shape[:3]
        d, k = self.embed_dim, self.conv_kernel
        w_oihw, bias = self._dynamic_weights_autograd(wavelengths, c)
        kk = k * k * c
        kpad = (kk + 63) // 64 * 64
        col = ops.im2col(img, c, k, k, k, 1, kpad)
        wp = ops.pack_conv_weight(w_oihw.detach().to(acc).contiguous(), dt, 0, kpad)
        patch = ops.conv2d_fwd([col], wp, d, 1, 1, 0, 0, bias=bias.detach().to(acc).contiguous())
        hp, wpx = patch.shape[1:3]
        p_tok = hp * wpx
        if p_tok + 1 != self.pos_embed.shape[1]:
            raise ValueError(f"image {hh}x{ww} gives {p_tok} patches but pos_embed has {self.pos_embed.shape[1] - 1}")
        tokens = ops.vit_assemble_tokens(patch.view(b, p_tok, d), self.pos_embed[0].detach().to(acc).contiguous(),
                                         self.cls_token.detach().to(acc).view(d).contiguous())
        n = p_tok + 1
        m = b * n
        heads = self.num_heads
        hd = d // heads
        lp = (n + 63) // 64 * 64
        stream = tokens.view(m, d)
        last = max(self.out_indices)
        blocks, feats = [], []
        for i, blk in enumerate(self.blocks[:last + 1]):
            s1, s2 = self._drop_path_factors(i, b, img.device, eng.training)
            x_in = stream
            a1, st1 = ops.layernorm_fwd(stream, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, dt, True)
            rc_qkv, act_a1 = self._lin_train(eng, a1, blk.attn.qkv)
            qkv = rc_qkv.x.view(m, 3 * d)
            q4 = qkv.view(b, 1, n, 3 * d)
            scores = torch.empty((b, 1, n, heads * lp), dtype=dt, device=img.device)
            ops.conv2d_fwd([q4[..., 0:hd]], qkv[:, d:d + hd], lp, 1, 1, 0, 0, out=scores[..., 0:lp], w_rows_per_img=n,
                           groups=(heads, hd, hd, lp))
            p4 = ops.softmax_fwd(scores.view(b, n, heads, lp), hd ** -0.5, n).view(b, 1, n, heads * lp)
            o = torch.empty((b, 1, n, d), dtype=dt, device=img.device)
            ops.conv2d_fwd([p4[..., 0:lp]], qkv[:, 2 * d:2 * d + hd], hd, 1, 1, 0, 0, out=o[..., 0:hd], w_rows_per_img=n,
                           w_mn_major=True, groups=(heads, lp, hd, hd))
            rc_proj, act_o = self._lin_train(eng, o.view(m, d), blk.attn.proj)
            stream = ops.layerscale_add(stream, rc_proj.x.view(m, d), blk.ls1.gamma.detach().to(acc), s1, n)
            x_mid = stream
            a2, st2 = ops.layernorm_fwd(stream, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, dt, True)
            rc_fc1, act_a2 = self._lin_train(eng, a2, blk.mlp.fc1)
            f = ops.gelu_fwd(rc_fc1.x)
            rc_fc2, act_f = self._lin_train(eng, f.view(m, -1), blk.mlp.fc2)
            stream = ops.layerscale_add(stream, rc_fc2.x.view(m, d), blk.ls2.gamma.detach().to(acc), s2, n)
            blocks.append(_SavedBlock(blk=blk, x_in=x_in, st1=st1, rc_qkv=rc_qkv, act_a1=act_a1, p4=p4, rc_proj=rc_proj,
                                      act_o=act_o, s1=s1, x_mid=x_mid, st2=st2, rc_fc1=rc_fc1, act_a2=act_a2,
                                      rc_fc2=rc_fc2, act_f=act_f, s2=s2))
            if i in self.out_indices:
                feats.append(Act(ops.vit_extract_feature(stream.view(b, n, d), dt).view(b, hp, wpx, d)))
        self._saved = _SavedBlock(blocks=blocks, feats=feats, col=col, kpad=kpad, w_oihw=w_oihw, bias=bias, b=b, n=n, c=c,
                                  lp=lp, hp=hp, wpx=wpx)
        return feats

    @staticmethod
    def _take(act: Act) -> torch.Tensor:
        assert len(act.gsrcs) == 1 and act.gsrcs[0][1] == 0
        g = act.gsrcs[0][0]
        act.gsrcs.clear()
        return g

    def _ln_grads(self, eng: Engine, ln: nn.LayerNorm):
        need = ln.weight.requires_grad or ln.bias.requires_grad
        return torch.zeros((2, ln.weight.numel()), dtype=eng.acc_dtype, device=ln.weight.device) if need else None

    def _store_ln(self, eng: Engine, ln: nn.LayerNorm, pg) -> None:
        if pg is None:
            return
        if ln.weight.requires_grad:
            eng.grad_buffer(ln.weight, False).copy_(pg[0])
        if ln.bias.requires_grad:
            eng.grad_buffer(ln.bias, False).copy_(pg[1])

    def _layerscale_bwd(self, eng: Engine, ls: _LayerScale, g: torch.Tensor, u: torch.Tensor, sscale, n: int):
        dg = torch.zeros(ls.gamma.numel(), dtype=eng.acc_dtype, device=g.device) if ls.gamma.requires_grad else None
        du = ops.layerscale_bwd(g, u, ls.gamma.detach().to(eng.acc_dtype), dg, sscale, n)
        if dg is not None:
            eng.grad_buffer(ls.gamma, False).copy_(dg)
        return du

    def backward(self, eng: Engine) -> None:
        """Back-propagates the gradients registered on the maps returned by `run_train` through the blocks, the token
        glue, the dynamic patch embedding and (torch autograd) the weight generator; parameter gradients go to the
        engine's gradient buffers."""
        S = self._saved
        dt, acc = eng.dtype, eng.acc_dtype
        b, n, d, lp = S.b, S.n, self.embed_dim, S.lp
        m, heads = b * n, self.num_heads
        hd = d // heads
        taps = {idx: f for idx, f in zip(sorted(self.out_indices), S.feats)}
        g = None  # fp32 gradient of the residual stream, (m, d)
        for i in range(len(S.blocks) - 1, -1, -1):
            sv = S.blocks[i]
            blk = sv.blk
            if i in taps and taps[i].gsrcs:
                dfeat = eng.collect_grad(taps[i])
                g3 = ops.vit_feature_grad(dfeat.reshape(b, n - 1, d), g.view(b, n, d) if g is not None else None)
                g = g3.view(m, d)
            if g is None:
                continue
            # ---- x = x_mid + drop_path(ls2(fc2(gelu(fc1(norm2(x_mid))))))
            du2 = self._layerscale_bwd(eng, blk.ls2, g, sv.rc_fc2.x.view(m, d), sv.s2, n)
            eng.conv_backward(sv.rc_fc2, du2.view(1, 1, m, d))
            dpre = ops.gelu_bwd(self._take(sv.act_f), sv.rc_fc1.x)
            eng.conv_backward(sv.rc_fc1, dpre)
            pg = self._ln_grads(eng, blk.norm2)
            g, _ = ops.layernorm_bwd(self._take(sv.act_a2).view(m, d), sv.x_mid, sv.st2, blk.norm2.weight, add=g, want32=True,
                                     pgrads=pg)
            self._store_ln(eng, blk.norm2, pg)
            # ---- x_mid = x_in + drop_path(ls1(proj(softmax(q k^T / sqrt(hd)) v)))
            du1 = self._layerscale_bwd(eng, blk.ls1, g, sv.rc_proj.x.view(m, d), sv.s1, n)
            eng.conv_backward(sv.rc_proj, du1.view(1, 1, m, d))
            do4 = self._take(sv.act_o).view(b, 1, n, d)
            qkv = sv.rc_qkv.x.view(m, 3 * d)
            q4 = qkv.view(b, 1, n, 3 * d)
            dp = torch.empty((b, 1, n, heads * lp), dtype=dt, device=g.device)
            ops.conv2d_fwd([do4[..., 0:hd]], qkv[:, 2 * d:2 * d + hd], lp, 1, 1, 0, 0, out=dp[..., 0:lp], w_rows_per_img=n,
                           groups=(heads, hd, hd, lp))
            ds4 = ops.softmax_bwd(sv.p4.view(b, n, heads, lp), dp.view(b, n, heads, lp), hd ** -0.5, n).view(b, 1, n, heads * lp)
            dqkv = torch.empty((b, 1, n, 3 * d), dtype=dt, device=g.device)
            ops.conv2d_fwd([ds4[..., 0:lp]], qkv[:, d:d + hd], hd, 1, 1, 0, 0, out=dqkv[..., 0:hd], w_rows_per_img=n,
                           w_mn_major=True, groups=(heads, lp, hd, hd))
            dkv32 = torch.zeros((b, lp, 2 * d), dtype=acc, device=g.device)
            # dK[b] = dS^T q,  dV[b] = P^T dO   (one independent product per image and head; key rows >= n stay zero)
            if ops.option("attn_wgrad_grouped"):
                ops.conv2d_wgrad([q4[..., 0:hd]], ds4[..., 0:lp], 1, 1, 0, 0, dkv32[:, :, 0:hd], groups=(heads, hd, lp, hd))
                ops.conv2d_wgrad([do4[..., 0:hd]], sv.p4[..., 0:lp], 1, 1, 0, 0, dkv32[:, :, d:d + hd], groups=(heads, hd, lp, hd))
            else:
                for h_ in range(heads):
                    ops.conv2d_wgrad([q4[..., h_ * hd:(h_ + 1) * hd]], ds4[..., h_ * lp:(h_ + 1) * lp], 1, 1, 0, 0,
                                     dkv32[:, :, h_ * hd:(h_ + 1) * hd])
                    ops.conv2d_wgrad([do4[..., h_ * hd:(h_ + 1) * hd]], sv.p4[..., h_ * lp:(h_ + 1) * lp], 1, 1, 0, 0,
                                     dkv32[:, :, d + h_ * hd:d + (h_ + 1) * hd])
            dqkv.view(b, n, 3 * d)[:, :, d:].copy_(dkv32[:, :n])  # cast + column placement (host-side glue)
            eng.conv_backward(sv.rc_qkv, dqkv.view(1, 1, m, 3 * d))
            pg = self._ln_grads(eng, blk.norm1)
            g, _ = ops.layernorm_bwd(self._take(sv.act_a1).view(m, d), sv.x_in, sv.st1, blk.norm1.weight, add=g, want32=True,
                                     pgrads=pg)
            self._store_ln(eng, blk.norm1, pg)
        self._saved = None
        if g is None:
            return
        # ---- tokens = [cls ; patch + pos]: d(cls) = sum_b g[b][0]; d(patch) = g[:, 1:]
        g3 = g.view(b, n, d)
        if self.cls_token.requires_grad:
            eng.grad_buffer(self.cls_token, False).copy_(g3[:, 0].sum(0).view(1, 1, d))
        gen = self._generator_params()
        if not gen:
            return
        dpatch = ops.vit_extract_feature(g3, dt).view(b, S.hp, S.wpx, d)
        dw = torch.zeros((d, S.kpad), dtype=acc, device=g.device)
        ops.conv2d_wgrad([S.col], dpatch, 1, 1, 0, 0, dw)
        k = self.conv_kernel
        dw_oihw = torch.empty((d, S.c, k, k), dtype=acc, device=g.device)
        ops.unpack_conv_wgrad(dw, dw_oihw, S.kpad)
        sums = torch.empty(2 * d, dtype=acc, device=g.device)
        ops.bn_stats(dpatch, sums)  # column sums: the bias gradient
        grads = torch.autograd.grad([S.w_oihw, S.bias], gen, [dw_oihw.to(S.w_oihw.dtype), sums[:d].to(S.bias.dtype)],
                                    allow_unused=True)
        for p_, gp in zip(gen, grads):
            if gp is not None:
                eng.grad_buffer(p_, False).copy_(gp)

    def forward(self, x: torch.Tensor, wavelengths: torch.Tensor) -> list[torch.Tensor]:
        """returns the reference's format: list of (B, D, h, w) tensors (channels-last memory)"""
        return [f.permute(0, 3, 1, 2) for f in self.forward_features(x, wavelengths)]


def create_dofa_base(img_size=512, out_indices=(4, 6, 10, 11), *, pretrained: bool = False, **kw) -> DOFAv2:
    return DOFAv2("dofa_base", img_size, 14, 768, 12, 12, out_indices=list(out_indices), pretrained=pretrained, **kw)


def create_dofa_large(img_size=512, out_indices=(5, 11, 17, 23), *, pretrained: bool = False, **kw) -> DOFAv2:
    return DOFAv2("dofa_large", img_size, 14, 1024, 24, 16, out_indices=list(out_indices), pretrained=pretrained, **kw)


class DOFASegmentationModel(UperNetSegmentor):
    """encoder (frozen DOFA ViT) + MultiLevelNeck + UperNetDecoder + SegmentationHead + FCNHead aux head."""

    def __init__(self, encoder: str = "dofa_base", image_size=(512, 512), freeze_layers: list[str] | None = None,
                 num_classes: int = 1, *, pretrained: bool = False, compute_dtype: torch.dtype = torch.bfloat16,
                 aux_dropout_ratio: float = 0.0, drop_path_rate: float = 0.1) -> None:
        if encoder not in ("dofa_base", "dofa_large"):
            raise ValueError(f"Invalid encoder: {encoder}")
        dim = 768 if encoder == "dofa_base" else 1024
        super().__init__(dim, 256, num_classes, compute_dtype, aux_dropout_ratio)
        make = create_dofa_base if encoder == "dofa_base" else create_dofa_large
        self.encoder = make(img_size=image_size, pretrained=pretrained, compute_dtype=compute_dtype,
                            drop_path_rate=drop_path_rate)
        if freeze_layers:
            for n_, p in self.named_parameters():
                if any(layer in n_ for layer in freeze_layers):
                    p.requires_grad = False

    def forward(self, x: torch.Tensor, wavelengths: torch.Tensor) -> SegmentationOutput:  # type: ignore[override]
        image_size = tuple(x.shape[2:])
        if torch.is_grad_enabled() and self.training and self.encoder.trainable():
            ops.require_cuda(x, "gdl_b200.DOFASegmentationModel")
            params = [p for p in self.parameters() if p.requires_grad]
            out, aux = _DofaSegFn.apply(self, x, wavelengths, *params)
            return SegmentationOutput(out, aux)
        feats = self.encoder.forward_features(x, wavelengths)  # NHWC 16-bit, no gradient (frozen encoder)
        head_params = [p for n_, p in self.named_parameters() if not n_.startswith("encoder.")]
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in head_params):
            nchw = [f.permute(0, 3, 1, 2) for f in feats]
            out, aux = _UperNetFn.apply(self, image_size, len(nchw), *nchw, *head_params)
            return SegmentationOutput(out, aux)
        with torch.no_grad():
            eng = Engine(self.compute_dtype, training=False, wcache=self._wcache)
            o, a = self.run(eng, [Act(f, needs_grad=False) for f in feats], image_size)
            self._saved = None
        return SegmentationOutput(o.permute(0, 3, 1, 2), a.permute(0, 3, 1, 2))

    def fused_train(self, eng: Engine, x16: torch.Tensor, c: int, target: torch.Tensor, spec,
                    grad_scale: torch.Tensor | None = None) -> torch.Tensor:
        """FusedTrainer hook: normalised NHWC tiles -> loss = L(out) + 0.4 L(aux) (segmentation_dofa.py:226-228) with
        the gradients of the trainable half left in the engine's destination buffers.  Needs `self.wavelengths`."""
        train_enc = self.encoder.trainable()
        if train_enc:
            if x16.is_cuda and torch.cuda.is_current_stream_capturing():
                raise NotImplementedError("CUDA-graph capture of a step with a trainable DOFA encoder: the weight generator is "
                                          "differentiated by torch autograd (another thread); use cuda_graph=False")
            feats = self.encoder.run_train(eng, x16, c, self.wavelengths)
        else:
            feats = [Act(f, needs_grad=False) for f in self.encoder._features_nhwc(x16, c, self.wavelengths)]
        image_size = tuple(x16.shape[1:3])
        fused = bool(ops.option("fused_head"))
        o, a = self.run(eng, feats, image_size, upsample=not fused)
        if getattr(self, "_aux_w", None) is None or self._aux_w.device != o.device:
            self._aux_w = torch.full((1,), 0.4, dtype=torch.float32, device=o.device)
        # grad_scale: the trainer's fp16 loss scale S (device scalar): d(out) *= S, d(aux) *= 0.4 S
        aux_scale = self._aux_w if grad_scale is None else self._aux_w * grad_scale.to(self._aux_w.dtype)
        if fused:
            # the two bilinear resizes to the image size (models/segmentation/dofa.py:90-105) are fused with the loss:
            # neither (N,H,W,K) map nor its gradient is materialised
            co, _ = ops.upsample_ce_fwd(o, target, spec)
            ca, _ = ops.upsample_ce_fwd(a, target, spec)
            kp = (o.shape[3] + 15) // 16 * 16
            d_o = torch.zeros((*o.shape[:3], kp), dtype=eng.dtype, device=o.device)
            d_a = torch.zeros((*a.shape[:3], kp), dtype=eng.dtype, device=a.device)
            ops.upsample_ce_bwd(o, target, spec, co, grad_scale, d_o)
            ops.upsample_ce_bwd(a, target, spec, ca, aux_scale, d_a)
            self.backward(eng, d_o, d_a, lowres16=True)
        else:
            co, _ = ops.seg_loss_fwd(o, target, spec)
            ca, _ = ops.seg_loss_fwd(a, target, spec)
            d_o, d_a = torch.empty_like(o), torch.empty_like(a)
            ops.seg_loss_bwd(o, target, spec, co, grad_scale, d_o)
            ops.seg_loss_bwd(a, target, spec, ca, aux_scale, d_a)
            self.backward(eng, d_o, d_a)
        if train_enc:
            self.encoder.backward(eng)
        return co[0] + 0.4 * ca[0]

    def _feat(self, f: torch.Tensor, needs_grad: bool) -> Act:
        # features coming from the B200 encoder are already NHWC 16-bit (logical NCHW view): no copy
        nhwc = f.permute(0, 2, 3, 1)
        if nhwc.dtype == self.compute_dtype and nhwc.is_contiguous():
            return Act(nhwc, needs_grad=needs_grad)
        return super()._feat(f, needs_grad)


class _DofaSegFn(torch.autograd.Function):
    """encoder (trainable) + neck + UperNet + heads as ONE autograd node: the hand-written backward of both halves runs
    on one engine, the map gradients never leave the 16-bit NHWC layout."""

    @staticmethod
    def forward(ctx, model: DOFASegmentationModel, x: torch.Tensor, wavelengths: torch.Tensor, *params: torch.Tensor):
        eng = Engine(model.compute_dtype, training=True, wcache=model._wcache, sync_bn_group=model.sync_bn_group,
                     acc_dtype=getattr(model, "acc_dtype", torch.float32))
        c = x.shape[1]
        img = ops.normalize_to_nhwc(x.contiguous().to(eng.acc_dtype), True, model.compute_dtype, (c + 7) // 8 * 8)
        feats = model.encoder.run_train(eng, img, c, wavelengths)
        o, a = model.run(eng, feats, tuple(x.shape[2:]))
        ctx.eng, ctx.model, ctx.params = eng, model, params
        return o.permute(0, 3, 1, 2), a.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, d_out: torch.Tensor, d_aux: torch.Tensor):
        eng: Engine = ctx.eng

        def nhwc(dd):
            return None if dd is None else dd.permute(0, 2, 3, 1).contiguous().to(eng.acc_dtype)
        ctx.model.backward(eng, nhwc(d_out), nhwc(d_aux))
        ctx.model.encoder.backward(eng)
        grads = tuple(eng.param_grads.get(id(p)) for p in ctx.params)
        ctx.eng = None
        return (None, None, None, *grads)
