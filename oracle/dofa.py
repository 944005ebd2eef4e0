"""ORACLE (test infrastructure — never imported by the product path).

Functional CPU/fp32 restatement of the reference's DOFA-v2 encoder forward
(geo_deep_learning/models/encoders/dofa_v2.py): wavelength sin/cos embedding (:9-35), FCResLayer (:38-56),
TransformerWeightGenerator (:59-106; torch's post-norm nn.TransformerEncoderLayer, 4 heads, GELU, ff 2048),
dynamic 14x14/stride-14/pad-1 patch embedding with the 0.01 scaler (:148-181), fixed sin/cos pos-embed on the
patch tokens only, cls token prepended without pos-embed (:445-455), 12 ViT blocks with feature taps (:461-487;
the final norm is never applied when depth-1 is tapped — reference quirk kept).

The ViT block is `timm.models.vision_transformer.Block` (timm 1.0.24, un-vendored, uv.lock.cu128:3457, absent here):
restated from its published definition — x += ls1(attn(norm1 x)); x += ls2(mlp(norm2 x)); attn = fused qkv Linear
(bias) + scaled-dot-product attention (scale d^-1/2) + proj; ls* = per-channel gamma; mlp = fc1 -> exact GELU -> fc2.

PARITY STATUS: embedding / weight generator / token glue / taps are **pinned** against the reference's own DOFAv2
module (imported with oracle/ref_shims.py, which supplies this same Block restatement) in tests/test_oracle_cpu.py;
the Block's internals are **unpinned** (no timm in this environment, no reference test).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

OUT_INDICES_BASE = (4, 6, 10, 11)  # create_dofa_base (dofa_v2.py:531)


def position_embedding(embed_dim, pos):
    omega = torch.arange(embed_dim // 2, dtype=torch.float32, device=pos.device) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = torch.einsum("m,d->md", pos.reshape(-1), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1)


def sincos_2d(embed_dim, grid):
    """DOFAv2.get_2d_sincos_pos_embed(cls_token=True): (1 + grid^2, D), row 0 zeros"""
    gh, gw = torch.meshgrid(torch.arange(grid), torch.arange(grid), indexing="ij")

    def one(d, p):
        om = 1.0 / 10000 ** (torch.arange(d // 2, dtype=torch.float32) / (d / 2.0))
        o = torch.einsum("m,d->md", p.reshape(-1).float(), om)
        return torch.cat([torch.sin(o), torch.cos(o)], dim=1)
    pe = torch.cat([one(embed_dim // 2, gh), one(embed_dim // 2, gw)], dim=1)
    return torch.cat([torch.zeros(1, embed_dim), pe], dim=0)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def weight_generator(sd, waves, p="patch_embed.weight_generator."):
    """TransformerWeightGenerator.forward: (C,128) -> weights (C, 14*14*D), bias (D,)"""
    x = torch.cat([sd[p + "weight_tokens"], waves, sd[p + "bias_token"]], dim=0)  # (L, 128), unbatched seq-first
    lp = p + "transformer_encoder.layers.0."
    e, heads = x.shape[1], 4
    qkv = F.linear(x, sd[lp + "self_attn.in_proj_weight"], sd[lp + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(e, dim=1)
    d = e // heads
    q, k, v = (t.view(-1, heads, d).transpose(0, 1) for t in (q, k, v))
    a = torch.softmax((q @ k.transpose(1, 2)) * d ** -0.5, dim=-1) @ v
    a = _lin(a.transpose(0, 1).reshape(-1, e), sd, lp + "self_attn.out_proj")
    x = F.layer_norm(x + a, (e,), sd[lp + "norm1.weight"], sd[lp + "norm1.bias"], 1e-5)
    f = _lin(F.gelu(_lin(x, sd, lp + "linear1")), sd, lp + "linear2")
    x = F.layer_norm(x + f, (e,), sd[lp + "norm2.weight"], sd[lp + "norm2.bias"], 1e-5)
    wt = 128
    return _lin(x[wt:-1] + waves, sd, p + "fc_weight"), _lin(x[-1], sd, p + "fc_bias")


def patch_embed(sd, img, wavelengths, embed_dim, k=14, convert_to_16=False):
    c = img.shape[1]
    waves = position_embedding(128, wavelengths * 1000)
    y = F.relu(_lin(F.relu(_lin(waves, sd, "patch_embed.fclayer.w1")), sd, "patch_embed.fclayer.w2"))
    waves = waves + y
    w, b = weight_generator(sd, waves)
    w = w.view(c, k, k, embed_dim).permute(3, 0, 1, 2) * 0.01
    if convert_to_16:  # dofa_v2.py:168-176
        w = F.interpolate(w, size=(16, 16), mode="bicubic", align_corners=False)
        k = 16
    x = F.conv2d(img, w, b.view(embed_dim) * 0.01, stride=k, padding=1)
    return x.flatten(2).transpose(1, 2)


def vit_block(sd, x, p, heads, drop_path=None):
    """timm Block; `drop_path` = ((B,), (B,)) per-sample factors keep_mask / keep_prob of the two branches (timm's
    DropPath in train mode, with the random draw supplied by the caller) or None (eval / rate 0)."""
    b, n, c = x.shape
    d = c // heads
    h = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    qkv = _lin(h, sd, p + "attn.qkv").view(b, n, 3, heads, d).permute(2, 0, 3, 1, 4)
    a = torch.softmax((qkv[0] @ qkv[1].transpose(-2, -1)) * d ** -0.5, dim=-1) @ qkv[2]
    a = _lin(a.transpose(1, 2).reshape(b, n, c), sd, p + "attn.proj")
    s1 = s2 = 1.0
    if drop_path is not None and drop_path[0] is not None:
        s1, s2 = drop_path[0].to(x.dtype).view(b, 1, 1), drop_path[1].to(x.dtype).view(b, 1, 1)
    x = x + s1 * (a * sd[p + "ls1.gamma"])
    h = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    m = _lin(F.gelu(_lin(h, sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    return x + s2 * (m * sd[p + "ls2.gamma"])


def dofa_forward(sd, img, wavelengths, embed_dim=768, depth=12, heads=12, out_indices=OUT_INDICES_BASE, drop_path=None,
                 convert_to_16=False):
    """img (B,C,H,W), wavelengths (C,) in micrometres -> list of (B, D, H/14', W/14') maps"""
    x = patch_embed(sd, img, wavelengths, embed_dim, convert_to_16=convert_to_16) + sd["pos_embed"][:, 1:, :]
    x = torch.cat([sd["cls_token"].expand(x.shape[0], -1, -1), x], dim=1)
    feats = []
    for i in range(depth):
        x = vit_block(sd, x, f"blocks.{i}.", heads, None if drop_path is None else drop_path[i])
        if i in out_indices:
            f = x[:, 1:, :]
            b, n, c = f.shape
            s = int(n ** 0.5)
            feats.append(f.reshape(b, s, s, c).permute(0, 3, 1, 2))
    return feats


def init_state_dict(embed_dim=768, depth=12, img=512, seed=0, ls_init=1e-5):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(p, cin, cout, std=None):
        sd[p + ".weight"] = torch.randn(cout, cin, generator=g) * (std or (2.0 / (cin + cout)) ** 0.5)
        sd[p + ".bias"] = torch.full((cout,), 0.01) + 0.01 * torch.randn(cout, generator=g)

    def ln(p, c):
        sd[p + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[p + ".bias"] = 0.05 * torch.randn(c, generator=g)

    pe = "patch_embed."
    wg = pe + "weight_generator."
    sd[wg + "weight_tokens"] = torch.randn(128, 128, generator=g) * 0.02
    sd[wg + "bias_token"] = torch.randn(1, 128, generator=g) * 0.02
    lp = wg + "transformer_encoder.layers.0."
    sd[lp + "self_attn.in_proj_weight"] = torch.randn(384, 128, generator=g) * (2.0 / 512) ** 0.5
    sd[lp + "self_attn.in_proj_bias"] = 0.01 * torch.randn(384, generator=g)
    lin(lp + "self_attn.out_proj", 128, 128)
    lin(lp + "linear1", 128, 2048)
    lin(lp + "linear2", 2048, 128)
    ln(lp + "norm1", 128)
    ln(lp + "norm2", 128)
    lin(wg + "fc_weight", 128, 14 * 14 * embed_dim)
    lin(wg + "fc_bias", 128, embed_dim)
    lin(pe + "fclayer.w1", 128, 128)
    lin(pe + "fclayer.w2", 128, 128)
    n = (img // 14) ** 2
    sd["pos_embed"] = sincos_2d(embed_dim, int(n ** 0.5)).unsqueeze(0)
    sd["cls_token"] = torch.randn(1, 1, embed_dim, generator=g) * 0.02
    for i in range(depth):
        p = f"blocks.{i}."
        ln(p + "norm1", embed_dim)
        lin(p + "attn.qkv", embed_dim, 3 * embed_dim, 0.02)
        lin(p + "attn.proj", embed_dim, embed_dim, 0.02)
        sd[p + "ls1.gamma"] = torch.full((embed_dim,), ls_init) * (1 + 0.1 * torch.randn(embed_dim, generator=g))
        ln(p + "norm2", embed_dim)
        lin(p + "mlp.fc1", embed_dim, 4 * embed_dim, 0.02)
        lin(p + "mlp.fc2", 4 * embed_dim, embed_dim, 0.02)
        sd[p + "ls2.gamma"] = torch.full((embed_dim,), ls_init) * (1 + 0.1 * torch.randn(embed_dim, generator=g))
    ln("norm", embed_dim)
    return sd
