"""ORACLE (test infrastructure — never imported by the product path).

Functional CPU/fp32 restatement of the reference's SegFormer path:
  encoder  geo_deep_learning/models/encoders/mix_transformer.py  (OverlapPatchEmbed :224-276,
           Attention :66-157, Mlp/DWConv :17-63,533-546, Block :160-221, stage loop :489-526)
  decoder  geo_deep_learning/models/decoders/segformer_mlp.py:8-130
  model    geo_deep_learning/models/segmentation/segformer.py:47-57 (final bilinear x4)

It is written over a plain `state_dict` (same keys as `SegFormerSegmentationModel`) instead of
nn.Modules, so it travels to the GPU box where /root/reference does not exist.

PARITY STATUS: **pinned** — tests/test_oracle_cpu.py compares it with the reference's own modules
(imported from /root/reference through a 20-line timm shim, oracle/ref_shims.py) when that tree is
present, and tests/golden/segformer_b0_golden.pt holds the reference's outputs (logits slice, loss,
gradient norms for seeded weights/inputs) produced by oracle/make_golden.py for where it is not.

Stochastic parts (DropPath, Dropout2d) are identity here: parity runs use eval mode or drop rates 0,
as SURVEY.md §5 prescribes; BatchNorm in `linear_fuse` follows `training`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# name: (embed_dims, heads, depths, decoder embedding dim)   [mix_transformer.py:599-708, segformer_mlp.py:40-44]
MIT_CFG = {
    "mit_b0": ((32, 64, 160, 256), (1, 2, 5, 8), (2, 2, 2, 2), 256),
    "mit_b1": ((64, 128, 320, 512), (1, 2, 5, 8), (2, 2, 2, 2), 256),
    "mit_b2": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 4, 6, 3), 768),
    "mit_b3": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 4, 18, 3), 768),
    "mit_b4": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 8, 27, 3), 768),
    "mit_b5": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 6, 40, 3), 768),
}
SR_RATIOS = (8, 4, 2, 1)
EPS_BLOCK = 1e-6  # norm_layer=partial(LayerNorm, eps=1e-6): block norm1/norm2 and the stage norms
EPS_PLAIN = 1e-5  # plain nn.LayerNorm: OverlapPatchEmbed.norm (:251) and Attention.norm (:100)


def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def attention(x, h, w, sd, p, heads, sr):
    b, n, c = x.shape
    d = c // heads
    q = _linear(x, sd, p + ".q").view(b, n, heads, d).transpose(1, 2)
    if sr > 1:
        xm = x.transpose(1, 2).reshape(b, c, h, w)
        xm = F.conv2d(xm, sd[p + ".sr.weight"], sd[p + ".sr.bias"], stride=sr)
        xr = _ln(xm.flatten(2).transpose(1, 2), sd, p + ".norm", EPS_PLAIN)
    else:
        xr = x
    kv = _linear(xr, sd, p + ".kv").view(b, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    a = torch.softmax((q @ k.transpose(-2, -1)) * d ** -0.5, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(b, n, c)
    return _linear(o, sd, p + ".proj")


def mix_ffn(x, h, w, sd, p):
    b, n, _ = x.shape
    y = _linear(x, sd, p + ".fc1")
    ch = y.shape[-1]
    ym = y.transpose(1, 2).reshape(b, ch, h, w)
    ym = F.conv2d(ym, sd[p + ".dwconv.dwconv.weight"], sd[p + ".dwconv.dwconv.bias"], padding=1, groups=ch)
    y = F.gelu(ym.flatten(2).transpose(1, 2))
    return _linear(y, sd, p + ".fc2")


def channel_position_encoding(n_channels, pos_dim, dtype=torch.float32):
    """DynamicChannelEmbed.get_position_encoding (mix_transformer.py:813-824): sinusoidal code of the band index"""
    positions = torch.arange(n_channels).float()
    dim_t = torch.arange(0, pos_dim, 2).float()
    inv_freq = 1.0 / (10000 ** (dim_t / pos_dim))
    pe = torch.zeros(n_channels, pos_dim)
    pe[:, 0::2] = torch.sin(positions.unsqueeze(1) * inv_freq)
    pe[:, 1::2] = torch.cos(positions.unsqueeze(1) * inv_freq)
    return pe.to(dtype)


def dynamic_channel_embed(sd, x, p="encoder.dynamic_patch_embed1"):
    """DynamicChannelEmbed.forward (mix_transformer.py:826-865): per-band 7x7/4 conv shared by all bands, bounded
    per-band channel weights from the band's position code, channel attention over the bands, projection + LayerNorm."""
    b, c, hh, ww = x.shape
    pos_dim = sd[p + ".weight_gen.0.weight"].shape[1]
    pe = channel_position_encoding(c, pos_dim, x.dtype).to(x.device)
    cw = torch.tanh(F.linear(F.relu(F.linear(pe, sd[p + ".weight_gen.0.weight"], sd[p + ".weight_gen.0.bias"])),
                             sd[p + ".weight_gen.2.weight"], sd[p + ".weight_gen.2.bias"]))            # (C, E)
    xc = F.conv2d(x.reshape(b * c, 1, hh, ww), sd[p + ".spatial_conv.weight"], sd[p + ".spatial_conv.bias"], stride=4, padding=3)
    e, h, w = xc.shape[1:]
    xw = xc.view(b, c, e, h * w) * cw.unsqueeze(0).unsqueeze(-1)
    cat = torch.cat([xw, pe.unsqueeze(0).unsqueeze(-1).expand(b, -1, -1, h * w)], dim=2)
    xa = cat.permute(0, 3, 1, 2).reshape(b * h * w, c, e + pos_dim).transpose(1, 2)
    sc = F.conv1d(F.relu(F.conv1d(xa, sd[p + ".channel_attention.0.weight"], sd[p + ".channel_attention.0.bias"])),
                  sd[p + ".channel_attention.2.weight"], sd[p + ".channel_attention.2.bias"])
    sc = sc.transpose(1, 2).reshape(b, h * w, c).permute(0, 2, 1)
    a = torch.softmax(sc, dim=1).unsqueeze(2)
    agg = (xw * a).sum(dim=1).transpose(1, 2)                                                          # (B, hw, E)
    out = _ln(_linear(agg, sd, p + ".proj"), sd, p + ".norm", EPS_PLAIN)
    return out, h, w


def encoder(sd, img, name, prefix="encoder.", drop_path=None):
    """`drop_path`: list over all blocks (stage-major, as mix_transformer.py:341-343 numbers its rates) of
    ((B,), (B,)) factors keep_mask / keep_prob for the attention and Mix-FFN branches — timm's DropPath in train mode with
    the random draw supplied by the caller — or None (eval / rate 0)."""
    dims, heads, depths, _ = MIT_CFG[name]
    x = img
    feats = []
    bi = 0
    dynamic = f"{prefix}dynamic_patch_embed1.proj.weight" in sd  # DynamicMixTransformer (mix_transformer.py:868-934)
    for s in range(4):
        pe = f"{prefix}patch_embed{s + 1}"
        k, stride = (7, 4) if s == 0 else (3, 2)
        if s == 0 and dynamic:
            t, h, w = dynamic_channel_embed(sd, x, f"{prefix}dynamic_patch_embed1")
            b, c = t.shape[0], t.shape[2]
        else:
            x = F.conv2d(x, sd[pe + ".proj.weight"], sd[pe + ".proj.bias"], stride=stride, padding=k // 2)
            b, c, h, w = x.shape
            t = _ln(x.flatten(2).transpose(1, 2), sd, pe + ".norm", EPS_PLAIN)
        for i in range(depths[s]):
            bp = f"{prefix}block{s + 1}.{i}"
            s1 = s2 = 1.0
            if drop_path is not None and drop_path[bi][0] is not None:
                s1, s2 = (d.to(t.dtype).view(-1, 1, 1) for d in drop_path[bi])
            bi += 1
            t = t + s1 * attention(_ln(t, sd, bp + ".norm1", EPS_BLOCK), h, w, sd, bp + ".attn", heads[s], SR_RATIOS[s])
            t = t + s2 * mix_ffn(_ln(t, sd, bp + ".norm2", EPS_BLOCK), h, w, sd, bp + ".mlp")
        t = _ln(t, sd, f"{prefix}norm{s + 1}", EPS_BLOCK)
        x = t.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        feats.append(x)
    return feats


def decoder(sd, feats, training, prefix="decoder.", bn_momentum=0.1, dropout_mask=None):
    """`dropout_mask` (B, emb) = keep / (1 - p): the draw of nn.Dropout2d before linear_pred (segformer_mlp.py:129)"""
    c1 = feats[0]
    size = c1.shape[2:]
    ups = []
    for lvl in (4, 3, 2, 1):
        f = feats[lvl - 1]
        b, c, h, w = f.shape
        y = _linear(f.flatten(2).transpose(1, 2), sd, f"{prefix}linear_c{lvl}.proj")
        y = y.transpose(1, 2).reshape(b, -1, h, w)
        if lvl != 1:
            y = F.interpolate(y, size=size, mode="bilinear", align_corners=False)
        ups.append(y)
    x = F.conv2d(torch.cat(ups, 1), sd[prefix + "linear_fuse.0.weight"])
    x = F.batch_norm(x, sd[prefix + "linear_fuse.1.running_mean"], sd[prefix + "linear_fuse.1.running_var"],
                     sd[prefix + "linear_fuse.1.weight"], sd[prefix + "linear_fuse.1.bias"], training, bn_momentum, 1e-5)
    x = F.relu(x)
    if dropout_mask is not None:
        x = x * dropout_mask.to(x.dtype)[:, :, None, None]
    return F.conv2d(x, sd[prefix + "linear_pred.weight"], sd[prefix + "linear_pred.bias"])


def segformer_forward(sd, img, name="mit_b2", training=False, drop_path=None, dropout_mask=None):
    """logits (B, K, H, W). `sd` holds tensors (optionally requiring grad) under the reference's keys."""
    y = decoder(sd, encoder(sd, img, name, drop_path=drop_path), training, dropout_mask=dropout_mask)
    return F.interpolate(y, size=img.shape[2:], mode="bilinear", align_corners=False)


def init_state_dict(name="mit_b2", in_channels=3, num_classes=5, seed=0):
    """Seeded weights with the reference's shapes/keys (values follow its init distributions loosely;
    parity tests copy the SAME tensors into the product, so only shapes/keys matter here)."""
    g = torch.Generator().manual_seed(seed)
    dims, heads, depths, emb = MIT_CFG[name]
    sd: dict[str, torch.Tensor] = {}

    def lin(p, cin, cout, bias=True):
        sd[p + ".weight"] = torch.randn(cout, cin, generator=g) * 0.02 * 2.5
        if bias:
            sd[p + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def ln(p, c):
        sd[p + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[p + ".bias"] = 0.05 * torch.randn(c, generator=g)

    def conv(p, cin, cout, k, groups=1, bias=True):
        fan_out = k * k * cout // groups
        sd[p + ".weight"] = torch.randn(cout, cin // groups, k, k, generator=g) * (2.0 / fan_out) ** 0.5
        if bias:
            sd[p + ".bias"] = torch.randn(cout, generator=g) * 0.02

    cin = in_channels
    for s in range(4):
        c = dims[s]
        pe = f"encoder.patch_embed{s + 1}"
        conv(pe + ".proj", cin, c, 7 if s == 0 else 3)
        ln(pe + ".norm", c)
        for i in range(depths[s]):
            bp = f"encoder.block{s + 1}.{i}"
            ln(bp + ".norm1", c)
            lin(bp + ".attn.q", c, c)
            lin(bp + ".attn.kv", c, 2 * c)
            lin(bp + ".attn.proj", c, c)
            if SR_RATIOS[s] > 1:
                conv(bp + ".attn.sr", c, c, SR_RATIOS[s])
                ln(bp + ".attn.norm", c)
            ln(bp + ".norm2", c)
            lin(bp + ".mlp.fc1", c, 4 * c)
            conv(bp + ".mlp.dwconv.dwconv", 4 * c, 4 * c, 3, groups=4 * c)
            lin(bp + ".mlp.fc2", 4 * c, c)
        ln(f"encoder.norm{s + 1}", c)
        cin = c
    for lvl in (4, 3, 2, 1):
        lin(f"decoder.linear_c{lvl}.proj", dims[lvl - 1], emb)
    sd["decoder.linear_fuse.0.weight"] = torch.randn(emb, 4 * emb, 1, 1, generator=g) * (1.0 / (4 * emb)) ** 0.5
    sd["decoder.linear_fuse.1.weight"] = 1 + 0.1 * torch.randn(emb, generator=g)
    sd["decoder.linear_fuse.1.bias"] = 0.05 * torch.randn(emb, generator=g)
    sd["decoder.linear_fuse.1.running_mean"] = torch.zeros(emb)
    sd["decoder.linear_fuse.1.running_var"] = torch.ones(emb)
    sd["decoder.linear_fuse.1.num_batches_tracked"] = torch.tensor(0)
    conv("decoder.linear_pred", emb, num_classes, 1)
    return sd


def init_dynamic_state_dict(name="mit_b0", num_classes=5, seed=0):
    """init_state_dict with stage 1's patch embedding replaced by seeded DynamicChannelEmbed tensors (reference keys)"""
    sd = init_state_dict(name, 3, num_classes, seed)
    sd = {k: v for k, v in sd.items() if not k.startswith("encoder.patch_embed1.")}
    e = MIT_CFG[name][0][0]
    g = torch.Generator().manual_seed(seed + 1000)
    p = "encoder.dynamic_patch_embed1."
    shapes = {"weight_gen.0.weight": (128, 128), "weight_gen.0.bias": (128,), "weight_gen.2.weight": (e, 128),
              "weight_gen.2.bias": (e,), "spatial_conv.weight": (e, 1, 7, 7), "spatial_conv.bias": (e,),
              "channel_attention.0.weight": (e // 2, e + 128, 1), "channel_attention.0.bias": (e // 2,),
              "channel_attention.2.weight": (1, e // 2, 1), "channel_attention.2.bias": (1,),
              "proj.weight": (e, e), "proj.bias": (e,), "norm.bias": (e,)}
    for k, shp in shapes.items():
        fan_in = shp[1] * (shp[2] if len(shp) > 2 else 1) * (shp[3] if len(shp) > 3 else 1) if len(shp) > 1 else 16
        sd[p + k] = torch.randn(shp, generator=g) * (1.0 / fan_in) ** 0.5
    sd[p + "norm.weight"] = 1 + 0.1 * torch.randn(e, generator=g)
    return sd
