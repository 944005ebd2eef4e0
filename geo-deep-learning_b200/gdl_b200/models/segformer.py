"""B200-native SegFormer (MixTransformer encoder + all-MLP decoder + bilinear x4 head).

Drop-in for `SegFormerSegmentationModel(encoder, in_channels, weights, freeze_layers, num_classes)`
(geo_deep_learning/models/segmentation/segformer.py:15-57): same constructor keywords, same
`state_dict` keys/shapes (`encoder.patch_embed1.proj.weight`, `encoder.block1.0.attn.q.weight`,
`decoder.linear_fuse.0.weight`, ...), same forward contract (float NCHW image -> (N, K, H, W) logits).
The nn.* sub-modules only hold parameters; arithmetic runs on libgdlb200.so:

  * tokens stay NHWC == (B, N, C) end to end (no NLC<->NCHW copies, mix_transformer.py:499,541-546);
  * every Linear / 1x1 conv / patch-embed / sr conv is the tcgen05 implicit-GEMM kernel (bias and the
    residual add fused in its epilogue); attention is q.k^T -> softmax kernel -> P.V with the k/v slices
    of the kv tensor used in place as batched GEMM operands (no per-head transposes or copies);
  * the residual stream is kept in fp32 (what torch.autocast does), LayerNorm emits the 16-bit operand
    of the next GEMM;
  * the decoder reads its 4 upsampled projections as a virtual concat (the 3072-channel torch.cat of
    segformer_mlp.py:127 is never materialised).

Stochastic layers: the reference trains MiT with DropPath (`drop_path_rate` 0.1, linearly increasing over the blocks,
mix_transformer.py:341-343,614-705) and the decoder with `nn.Dropout2d(0.1)` before `linear_pred` (segformer_mlp.py:73,129).
Both are constructor arguments here and default to 0 — the deterministic model parity is defined on; the
Lightning task mirror passes the reference's values.  With a non-zero rate the residual add of a block leaves the GEMM
epilogue: the branch output is scaled per sample (timm DropPath = keep mask / keep probability) by `layerscale_add`.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from ..engine import Act, BNParams, Engine, RawConv

MIT_CFG = {
    # name: (embed_dims, heads, depths, decoder embedding dim)
    "mit_b0": ((32, 64, 160, 256), (1, 2, 5, 8), (2, 2, 2, 2), 256),
    "mit_b1": ((64, 128, 320, 512), (1, 2, 5, 8), (2, 2, 2, 2), 256),
    "mit_b2": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 4, 6, 3), 768),
    "mit_b3": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 4, 18, 3), 768),
    "mit_b4": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 8, 27, 3), 768),
    "mit_b5": ((64, 128, 320, 512), (1, 2, 5, 8), (3, 6, 40, 3), 768),
}
SR_RATIOS = (8, 4, 2, 1)


def _init(m: nn.Module) -> None:  # MixVisionTransformer._init_weights (mix_transformer.py:425-439)
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02, a=-2.0, b=2.0)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


class _PatchEmbed(nn.Module):
    def __init__(self, k: int, stride: int, cin: int, dim: int) -> None:
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, k, stride=stride, padding=k // 2)
        self.norm = nn.LayerNorm(dim)  # eps 1e-5 (plain nn.LayerNorm, mix_transformer.py:251)


class _Attention(nn.Module):
    def __init__(self, dim: int, heads: int, sr: int) -> None:
        super().__init__()
        self.q = nn.Linear(dim, dim)
        self.kv = nn.Linear(dim, 2 * dim)
        self.proj = nn.Linear(dim, dim)
        self.sr_ratio, self.num_heads = sr, heads
        if sr > 1:
            self.sr = nn.Conv2d(dim, dim, sr, stride=sr)
            self.norm = nn.LayerNorm(dim)  # eps 1e-5 (mix_transformer.py:100)


class _DWConv(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)


class _Mlp(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.dwconv = _DWConv(4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int, sr: int) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads, sr)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim)


class _DynamicChannelEmbed(nn.Module):
    """parameter holder of DynamicChannelEmbed (mix_transformer.py:762-811): same sub-module names / shapes"""

    def __init__(self, embed_dim: int, hidden_dim: int = 128) -> None:
        super().__init__()
        self.pos_dim = hidden_dim
        self.weight_gen = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, embed_dim), nn.Tanh())
        self.spatial_conv = nn.Conv2d(1, embed_dim, kernel_size=7, stride=4, padding=3)
        self.channel_attention = nn.Sequential(nn.Conv1d(embed_dim + hidden_dim, embed_dim // 2, 1), nn.ReLU(),
                                               nn.Conv1d(embed_dim // 2, 1, 1))
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.norm = nn.LayerNorm(embed_dim)  # eps 1e-5


class MixTransformerEncoder(nn.Module):
    def __init__(self, name: str, in_channels: int, dynamic: bool = False) -> None:
        super().__init__()
        if name not in MIT_CFG:
            raise KeyError(f"Wrong encoder name `{name}`, supported encoders: {list(MIT_CFG)}")
        self.name = name
        dims, heads, depths, _ = MIT_CFG[name]
        cin = in_channels
        for s in range(4):
            if not (dynamic and s == 0):  # DynamicMixTransformer drops patch_embed1 (mix_transformer.py:886-893)
                setattr(self, f"patch_embed{s + 1}", _PatchEmbed(7 if s == 0 else 3, 4 if s == 0 else 2, cin, dims[s]))
            cin = dims[s]
        for s in range(4):
            setattr(self, f"block{s + 1}", nn.ModuleList(_Block(dims[s], heads[s], SR_RATIOS[s]) for _ in range(depths[s])))
            setattr(self, f"norm{s + 1}", nn.LayerNorm(dims[s], eps=1e-6))
        self.apply(_init)
        if dynamic:  # built outside MixVisionTransformer in the reference: torch's default initialisation
            self.dynamic_patch_embed1 = _DynamicChannelEmbed(dims[0])
        self.dynamic = dynamic
        self.out_channels = dims


class _MLPProj(nn.Module):
    def __init__(self, cin: int, emb: int) -> None:
        super().__init__()
        self.proj = nn.Linear(cin, emb)


class SegFormerDecoder(nn.Module):
    def __init__(self, name: str, num_classes: int) -> None:
        super().__init__()
        dims, _, _, emb = MIT_CFG[name]
        for lvl in (4, 3, 2, 1):
            setattr(self, f"linear_c{lvl}", _MLPProj(dims[lvl - 1], emb))
        self.linear_fuse = nn.Sequential(nn.Conv2d(4 * emb, emb, 1, bias=False), nn.BatchNorm2d(emb), nn.ReLU(inplace=True))
        self.linear_pred = nn.Conv2d(emb, num_classes, 1)
        self.embedding_dim = emb


class _Saved:
    """per-op saved tensors for the hand-written backward"""

    def __init__(self, **kw) -> None:
        self.__dict__.update(kw)


def _lin_shape(w: torch.Tensor) -> tuple:
    return (w.shape[0], w.shape[1], 1, 1)


class SegFormer(nn.Module):
    needs_bands = True  # run() takes the real band count (the image arrives channel-padded to 8)

    def __init__(self, encoder: str = "mit_b0", in_channels: int = 3, weights: str | None = None,
                 freeze_layers: list[str] | None = None, num_classes: int = 1, *, use_dynamic_encoder: bool = False,
                 compute_dtype: torch.dtype = torch.bfloat16, drop_path_rate: float = 0.0,
                 dropout_ratio: float = 0.0) -> None:
        super().__init__()
        if encoder in MIT_CFG:
            nblk = sum(MIT_CFG[encoder][2])
            self.drop_path_rates = [float(v) for v in torch.linspace(0, drop_path_rate, nblk)]
        self.dropout_ratio = dropout_ratio
        # test hooks: the draws supplied from outside — list over blocks of ((B,), (B,)) factors; (N, emb) Dropout2d mask
        self.drop_path_masks: list | None = None
        self.dropout_mask: torch.Tensor | None = None
        self._ones: dict = {}
        if weights is not None:
            raise ValueError("weights must be None: load pretrained tensors through load_state_dict")
        # use_dynamic_encoder: DynamicMixTransformer (mix_transformer.py:868-934) — stage 1's patch embedding is replaced by
        # the band-count-agnostic DynamicChannelEmbed; `in_channels` is then ignored, as in the reference
        self.encoder = MixTransformerEncoder(encoder, in_channels, dynamic=use_dynamic_encoder)
        self._dyn_buf: dict = {}
        self.decoder = SegFormerDecoder(encoder, num_classes)
        self.name = encoder
        self.num_classes = num_classes
        self.compute_dtype = compute_dtype
        self._wcache: dict = {}
        self.sync_bn_group = None
        self.last_engine: Engine | None = None
        if freeze_layers:
            # BaseSegmentationModel._freeze_layers (models/segmentation/base.py:40-44), called by the reference BEFORE the
            # decoder exists (models/segmentation/segformer.py:42-45): only encoder parameters can match — a pattern such as
            # "proj" or "norm" must not freeze decoder.linear_c*.proj / linear_fuse / linear_pred
            for n, p in self.encoder.named_parameters():
                if any(layer in f"encoder.{n}" for layer in freeze_layers):
                    p.requires_grad = False

    # ====================================================================================== forward
    def _linear(self, eng: Engine, x: torch.Tensor, lin: nn.Linear, *, residual=None, out_dtype=None, needs_grad=True):
        a = Act(x, needs_grad=needs_grad)
        rc = eng.conv_raw([a], lin.weight, 1, 0, bias=lin.bias, out_dtype=out_dtype, wshape=_lin_shape(lin.weight),
                          residual=residual)
        return rc, a

    def _drop_path_factors(self, eng: Engine, i: int, b: int, dev):
        """per-sample factors keep_mask / keep_prob of block i's two branches, or (None, None)"""
        if not eng.training:
            return None, None
        if self.drop_path_masks is not None:
            return self.drop_path_masks[i]
        rate = self.drop_path_rates[i]
        if rate <= 0.0:
            return None, None
        keep = 1.0 - rate
        draw = torch.bernoulli(torch.full((2, b), keep, dtype=torch.float32, device=dev)) / keep
        return draw[0].contiguous(), draw[1].contiguous()

    def _unit_gamma(self, c: int, like: torch.Tensor) -> torch.Tensor:
        key = (c, like.dtype, like.device)
        if key not in self._ones:
            self._ones[key] = torch.ones(c, dtype=like.dtype, device=like.device)
        return self._ones[key]

    def _branch_out(self, eng: Engine, y: torch.Tensor, lin: nn.Linear, stream: torch.Tensor, s):
        """stream + drop_path(lin(y)): fused in the GEMM epilogue when there is no DropPath factor"""
        if s is None:
            rc, act = self._linear(eng, y, lin, residual=stream, out_dtype=stream.dtype)
            return rc.x, rc, act
        rc, act = self._linear(eng, y, lin)
        b, h, w, c = stream.shape
        out = ops.layerscale_add(stream.view(-1, c), rc.x.view(-1, c), self._unit_gamma(c, stream), s, h * w)
        return out.view(stream.shape), rc, act

    def _branch_grad(self, rc, s, g16: torch.Tensor, g32: torch.Tensor) -> torch.Tensor:
        """gradient w.r.t. the branch output: the stream gradient, scaled per sample under DropPath"""
        if s is None:
            return g16
        b, h, w, c = g32.shape
        du = ops.layerscale_bwd(g32.view(-1, c), rc.x.view(-1, c), self._unit_gamma(c, g32), None, s, h * w)
        return du.view(g16.shape)

    def _attention_fwd(self, eng: Engine, a16: torch.Tensor, attn: _Attention, stream: torch.Tensor, s=None):
        """returns (new fp32 stream, saved).  a16 = LN1(stream) as 16-bit tokens (B,h,w,C)."""
        b, h, w, c = a16.shape
        heads = attn.num_heads
        d = c // heads
        n = h * w
        dt = eng.dtype
        rc_q, act_a = self._linear(eng, a16, attn.q)
        q = rc_q.x
        sv = _Saved(act_a=act_a, rc_q=rc_q, heads=heads, d=d)
        if attn.sr_ratio > 1:
            rc_sr = eng.conv_raw([act_a], attn.sr.weight, attn.sr_ratio, 0, bias=attn.sr.bias)
            kvin, st_sr = ops.layernorm_fwd(rc_sr.x, attn.norm.weight, attn.norm.bias, attn.norm.eps, dt, eng.training)
            sv.rc_sr, sv.st_sr = rc_sr, st_sr
        else:
            kvin = a16
        rc_kv, act_kvin = self._linear(eng, kvin, attn.kv)
        kv = rc_kv.x
        nk = kv.shape[1] * kv.shape[2]
        lp = (nk + 15) // 16 * 16
        kv2 = kv.view(b * nk, 2 * c)
        q4 = q.view(b, 1, n, c)
        if ops.option("sra_fused") and ops.sra_attention_supported(n, nk, d, save_p=eng.training):
            # ONE kernel: q.k^T in TMEM -> softmax in registers -> P~ through shared memory -> P.V (csrc/sra_attention.cu); the
            # normalised probabilities are written only when a backward will read them
            o3, p3 = ops.sra_attention_fwd(q.view(b, n, c), kv2, heads, nk, d ** -0.5, save_p=eng.training)
            o = o3.view(b, 1, n, c)
            p4 = p3.view(b, 1, n, heads * lp) if p3 is not None else None
        else:
            scores = torch.empty((b, 1, n, heads * lp), dtype=dt, device=q.device)
            # all heads in one grouped launch: head g reads q / k / v columns shifted by g*d, writes score columns by g*lp
            ops.conv2d_fwd([q4[..., 0:d]], kv2[:, 0:d], nk, 1, 1, 0, 0, out=scores[..., 0:nk], w_rows_per_img=nk,
                           groups=(heads, d, d, lp))
            p = ops.softmax_fwd(scores.view(b, n, heads, lp), d ** -0.5, nk)
            p4 = p.view(b, 1, n, heads * lp)
            o = torch.empty((b, 1, n, c), dtype=dt, device=q.device)
            ops.conv2d_fwd([p4[..., 0:lp]], kv2[:, c:c + d], d, 1, 1, 0, 0, out=o[..., 0:d], w_rows_per_img=nk,
                           w_mn_major=True, groups=(heads, lp, d, d))
        new_stream, rc_proj, act_o = self._branch_out(eng, o.view(b, h, w, c), attn.proj, stream, s)
        sv.__dict__.update(rc_kv=rc_kv, act_kvin=act_kvin, kv2=kv2, q4=q4, p4=p4, nk=nk, lp=lp, rc_proj=rc_proj,
                           act_o=act_o, s=s)
        return new_stream, sv

    def _mlp_fwd(self, eng: Engine, a16: torch.Tensor, mlp: _Mlp, stream: torch.Tensor, s=None):
        rc1, act_a = self._linear(eng, a16, mlp.fc1)
        dw = mlp.dwconv.dwconv
        y, pre = ops.dwconv3x3_gelu_fwd(rc1.x, dw.weight.detach().view(dw.weight.shape[0], 9), dw.bias)
        new_stream, rc2, act_y = self._branch_out(eng, y, mlp.fc2, stream, s)
        return new_stream, _Saved(rc1=rc1, act_a=act_a, pre=pre, rc2=rc2, act_y=act_y, s=s)

    # ====================================================================================== DynamicChannelEmbed
    def _dyn_dense(self, dyn: _DynamicChannelEmbed, c: int):
        """The three band-wise layers of DynamicChannelEmbed written as ordinary dense convolutions over the bands
        concatenated along the channel dim (block-diagonal weights), built from the parameters under autograd so that the
        gradients the wgrad kernels produce for the dense tensors flow back to the parameters (tiny tensors):
          Wd (C*E, C, 7, 7): band c -> its E maps, spatial_conv.weight * tanh-bounded channel weight cw[c]   (:826-835)
          W1 (C*E/2, C*E), b1: channel_attention[0] on [xw_c ; pos_enc_c] — the position half is a per-band bias (:836-851)
          W2 (16, C*E/2), b2: channel_attention[2], one logit per band (rows >= C zero)"""
        pdt = dyn.proj.weight.dtype
        dev = dyn.proj.weight.device
        e = dyn.proj.in_features
        with torch.enable_grad():
            pos = torch.arange(c, device=dev).float()
            inv = 1.0 / (10000 ** (torch.arange(0, dyn.pos_dim, 2, device=dev).float() / dyn.pos_dim))
            pe = torch.zeros(c, dyn.pos_dim, device=dev)
            pe[:, 0::2] = torch.sin(pos.unsqueeze(1) * inv)
            pe[:, 1::2] = torch.cos(pos.unsqueeze(1) * inv)
            pe = pe.to(pdt)
            cw = dyn.weight_gen(pe)                                                    # (C, E), tanh-bounded
            eye = torch.eye(c, dtype=pdt, device=dev)
            w7 = dyn.spatial_conv.weight[:, 0]                                          # (E, 7, 7)
            wd = torch.einsum("ce,ekl,cd->cedkl", cw, w7, eye).reshape(c * e, c, 7, 7)
            bd = (cw * dyn.spatial_conv.bias.unsqueeze(0)).reshape(c * e)
            ca0, ca2 = dyn.channel_attention[0], dyn.channel_attention[2]
            w1a, w1b = ca0.weight[:, :e, 0], ca0.weight[:, e:, 0]                       # (E/2, E), (E/2, pos_dim)
            w1 = torch.einsum("je,cd->cjde", w1a, eye).reshape(c * (e // 2), c * e)
            b1 = (pe @ w1b.t() + ca0.bias.unsqueeze(0)).reshape(c * (e // 2))
            w2 = torch.zeros(16, c * (e // 2), dtype=pdt, device=dev)
            w2 = torch.cat([torch.einsum("j,cd->cdj", ca2.weight[0, :, 0], eye).reshape(c, c * (e // 2)), w2[c:]], 0)
            b2 = torch.cat([ca2.bias.expand(c), torch.zeros(16 - c, dtype=pdt, device=dev)], 0)
            return {"wd": wd, "bd": bd, "w1": w1.view(c * (e // 2), c * e, 1, 1), "b1": b1,
                    "w2": w2.view(16, c * (e // 2), 1, 1), "b2": b2}

    def _dyn_embed_fwd(self, eng: Engine, x: Act, c: int):
        """x: NHWC 16-bit image, c real bands.  Returns (fp32 token stream (B,h,w,E), saved)."""
        dyn = self.encoder.dynamic_patch_embed1
        if c > 16:
            raise ValueError("DynamicChannelEmbed on the B200 kernels handles at most 16 bands")
        dt, acc = eng.dtype, eng.acc_dtype
        graph = self._dyn_dense(dyn, c)
        key = (c, graph["wd"].device, acc)
        buf = self._dyn_buf.get(key)
        if buf is None:  # persistent leaves: in-place refreshed each step, so the engine's packed-weight cache stays valid
            buf = self._dyn_buf[key] = {k: torch.empty(v.shape, dtype=acc, device=v.device).requires_grad_(True)
                                        for k, v in graph.items()}
        with torch.no_grad():
            for k, v in graph.items():
                buf[k].copy_(v.detach())
        e = dyn.proj.in_features
        rc_xw = eng.conv_raw([x], buf["wd"], 4, 3, bias=buf["bd"], wshape=(c * e, c, 7, 7))
        act_xw = Act(rc_xw.x)
        rc_hid = eng.conv_raw([act_xw], buf["w1"], 1, 0, bias=buf["b1"], relu=True)
        act_hid = Act(rc_hid.x)
        rc_sc = eng.conv_raw([act_hid], buf["w2"], 1, 0, bias=buf["b2"], out_dtype=acc)
        pooled, attn = ops.channel_pool_fwd(rc_xw.x, rc_sc.x, c, eng.training)
        rc_proj, act_pool = self._linear(eng, pooled, dyn.proj)
        stream, st = ops.layernorm_fwd(rc_proj.x, dyn.norm.weight, dyn.norm.bias, dyn.norm.eps, acc, eng.training)
        return stream, _Saved(dyn=dyn, c=c, graph=graph, buf=buf, rc_xw=rc_xw, act_xw=act_xw, rc_hid=rc_hid, act_hid=act_hid,
                              rc_sc=rc_sc, attn=attn, rc_proj=rc_proj, act_pool=act_pool, st=st)

    def _dyn_embed_bwd(self, eng: Engine, sv, gstream: torch.Tensor) -> None:
        dyn, c, dt = sv.dyn, sv.c, eng.dtype
        if gstream.is_cuda and torch.cuda.is_current_stream_capturing():
            raise NotImplementedError("CUDA-graph capture of a step through DynamicChannelEmbed: its band-weight generator is "
                                      "differentiated by torch autograd; use cuda_graph=False")
        pg = self._pgrads_ln(eng, dyn.norm)
        _, dproj = ops.layernorm_bwd(gstream, sv.rc_proj.x, sv.st, dyn.norm.weight, want32=False, dtype16=dt, pgrads=pg)
        self._store_ln_grads(eng, dyn.norm, pg)
        eng.conv_backward(sv.rc_proj, dproj)
        dxw1, dsc = ops.channel_pool_bwd(self._take(sv.act_pool).contiguous(), sv.rc_xw.x, sv.attn, c)
        eng.conv_backward(sv.rc_sc, dsc)
        dpre = ops.relu_bwd(self._take(sv.act_hid).contiguous(), sv.rc_hid.x)
        eng.conv_backward(sv.rc_hid, dpre)
        dxw2 = self._take(sv.act_xw)
        dxw = torch.empty(sv.rc_xw.x.shape, dtype=dt, device=dxw1.device)
        ops.grad_gather([(dxw1, 0), (dxw2, 0)], sv.rc_xw.x.shape, dt, g=dxw)
        eng.conv_backward(sv.rc_xw, dxw)
        names = [k for k in sv.graph if id(sv.buf[k]) in eng.param_grads]
        params = [p_ for n_, p_ in dyn.named_parameters() if p_.requires_grad and not n_.startswith(("proj.", "norm."))]
        if params and names:
            grads = torch.autograd.grad([sv.graph[k] for k in names], params,
                                        [eng.param_grads.pop(id(sv.buf[k])).view(sv.graph[k].shape).to(sv.graph[k].dtype)
                                         for k in names], allow_unused=True)
            for p_, g_ in zip(params, grads):
                if g_ is not None:
                    eng.grad_buffer(p_, False).copy_(g_)

    def run(self, eng: Engine, x: Act, bands: int | None = None, upsample: bool = True) -> torch.Tensor:
        """x: NHWC 16-bit image (channels possibly zero padded; `bands` = the real channel count, needed by the dynamic
        encoder).  Returns fp32 logits (N,H,W,K); with upsample=False the decoder's own (N,H/4,W/4,K) map, for the fused
        upsample + loss / argmax head (ops.upsample_ce_*, ops.upsample_argmax)."""
        enc, dec = self.encoder, self.decoder
        dims, heads, depths, emb = MIT_CFG[self.name]
        dt = eng.dtype
        acc = eng.acc_dtype
        _, hh, ww, _ = x.t.shape
        if hh % 32 or ww % 32:
            raise ValueError(f"input height/width ({hh},{ww}) must be divisible by 32")
        stages = []
        cur = x
        blk_index = 0
        for s in range(4):
            sv_dyn = pe = rc_pe = st_pe = None
            if s == 0 and enc.dynamic:
                if bands is None:
                    raise ValueError("the dynamic encoder needs the real band count of the (channel-padded) image")
                stream, sv_dyn = self._dyn_embed_fwd(eng, cur, bands)
            else:
                pe = getattr(enc, f"patch_embed{s + 1}")
                k, stride = pe.proj.kernel_size[0], pe.proj.stride[0]
                rc_pe = eng.conv_raw([cur], pe.proj.weight, stride, k // 2, bias=pe.proj.bias)
                stream, st_pe = ops.layernorm_fwd(rc_pe.x, pe.norm.weight, pe.norm.bias, pe.norm.eps, acc, eng.training)
            blocks = []
            for blk in getattr(enc, f"block{s + 1}"):
                dp1, dp2 = self._drop_path_factors(eng, blk_index, stream.shape[0], stream.device)
                blk_index += 1
                a1, st1 = ops.layernorm_fwd(stream, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, dt, eng.training)
                x_in = stream
                stream, sv_attn = self._attention_fwd(eng, a1, blk.attn, stream, dp1)
                a2, st2 = ops.layernorm_fwd(stream, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, dt, eng.training)
                x_mid = stream
                stream, sv_mlp = self._mlp_fwd(eng, a2, blk.mlp, stream, dp2)
                blocks.append(_Saved(blk=blk, x_in=x_in, st1=st1, attn=sv_attn, x_mid=x_mid, st2=st2, mlp=sv_mlp))
            nrm = getattr(enc, f"norm{s + 1}")
            feat, st_out = ops.layernorm_fwd(stream, nrm.weight, nrm.bias, nrm.eps, dt, eng.training)
            feat_act = Act(feat)
            stages.append(_Saved(pe=pe, rc_pe=rc_pe, st_pe=st_pe, blocks=blocks, norm=nrm, x_out=stream, st_out=st_out,
                                 feat=feat_act, src=cur, dyn=sv_dyn))
            cur = feat_act
        # ---- decoder (segformer_mlp.py:77-130)
        h1, w1 = stages[0].feat.t.shape[1:3]
        projs = None
        if ops.option("decoder_folded"):
            rc_fuse = self._decoder_folded_fwd(eng, stages)
        else:
            projs = []
            for lvl in (4, 3, 2, 1):
                f = stages[lvl - 1].feat
                lin = getattr(dec, f"linear_c{lvl}").proj
                rc = eng.conv_raw([f], lin.weight, 1, 0, bias=lin.bias, wshape=_lin_shape(lin.weight))
                up = Act(ops.bilinear_fwd(rc.x, h1, w1)) if lvl != 1 else Act(rc.x)
                projs.append(_Saved(lvl=lvl, rc=rc, up=up))
            rc_fuse = eng.conv_raw([p.up for p in projs], dec.linear_fuse[0].weight, 1, 0)
        bn = dec.linear_fuse[1]
        bn_state = eng.bn_prepare(rc_fuse, BNParams(bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                                    bn.num_batches_tracked, bn.eps, bn.momentum or 0.1))
        z = eng.bn_act(rc_fuse, bn_state, relu=True)  # registers its own backward on the tape
        z = eng.dropout2d(z, self.dropout_ratio, self.dropout_mask)  # nn.Dropout2d before linear_pred (train mode only)
        rc_pred = eng.conv_raw([z], dec.linear_pred.weight, 1, 0, bias=dec.linear_pred.bias, out_dtype=acc)
        logits = ops.bilinear_fwd(rc_pred.x, hh, ww) if upsample else rc_pred.x
        eng.named = {f"c{s + 1}": stages[s].feat for s in range(4)}
        # the saved activations belong to THIS forward (its engine), not to the module: two training forwards before a
        # backward, or an eval forward in between, must not overwrite each other's graph
        eng.saved_segformer = _Saved(stages=stages, projs=projs, rc_pred=rc_pred, z=z, h1=h1, w1=w1, hh=hh, ww=ww)
        return logits

    # ---- decoder with linear_fuse's 1x1 conv moved in front of the resizes (option "decoder_folded") -----------------------
    # segformer_mlp.py:77-128 computes  Z = W . concat_l(resize(P_l c_l + b_l))  with W = linear_fuse[0] (no bias) split into
    # four (emb x emb) blocks W_l, P_l / b_l = linear_c{l}.proj, resize = bilinear to the stride-4 grid (identity for l = 1).
    # The resize is linear with taps that sum to 1 and acts per channel, so it commutes with the channel mix:
    #     Z = sum_l resize((W_l P_l) c_l) + sum_l W_l b_l .
    # M_l = W_l P_l (emb x C_l) is formed from the fp32 masters every step (0.6 GFLOP for B2, parameter space); the big
    # products then run at each level's own resolution with K = C_l instead of one 4*emb -> emb product on the stride-4
    # grid (B2, 512^2 tiles: 50 instead of 1290 GFLOP forward per 16 tiles), and the three resized emb-channel maps are
    # never written: gdl_bilinear_sum_fwd reads the low-resolution products and writes Z once.  Same parameters, same
    # state_dict, same function; gradients by the chain rule:  dM_l = dY_l^T c_l,  dP_l = W_l^T dM_l,
    # dW_l = dM_l P_l^T + s b_l^T,  db_l = W_l^T s  with  dY_l = resize^T(dZ),  s = column sums of dZ.
    def _decoder_folded_fwd(self, eng: Engine, stages) -> RawConv:
        dec = self.decoder
        emb, acc = dec.embedding_dim, eng.acc_dtype
        fuse_w = dec.linear_fuse[0].weight
        wf = fuse_w.detach().view(emb, 4 * emb).to(acc)
        levels, cbias = [], None
        for slot, lvl in enumerate((4, 3, 2, 1)):  # concat order of the reference: [_c4, _c3, _c2, _c1]
            lin = getattr(dec, f"linear_c{lvl}").proj
            wl = wf[:, slot * emb:(slot + 1) * emb]
            m = torch.matmul(wl, lin.weight.detach().to(acc)).contiguous()
            cb = torch.mv(wl, lin.bias.detach().to(acc))
            cbias = cb if cbias is None else cbias + cb
            levels.append(_Saved(lvl=lvl, lin=lin, wl=wl, m=m, feat=stages[lvl - 1].feat))
        ys = {}
        for lv in levels:
            cin = lv.m.shape[1]
            wp = ops.pack_conv_weight(lv.m.view(emb, cin, 1, 1), eng.dtype, 0)
            ys[lv.lvl] = ops.conv2d_fwd([lv.feat.t], wp, emb, 1, 1, 0, 0, bias=cbias if lv.lvl == 1 else None)
        z = ops.bilinear_sum_fwd(ys[1], [ys[4], ys[3], ys[2]])
        return RawConv(z, [], fuse_w, 1, 0, wshape=tuple(fuse_w.shape),
                       custom_backward=(lambda dz: self._decoder_folded_bwd(eng, levels, dz)) if eng.training else None)

    def _decoder_folded_bwd(self, eng: Engine, levels, dz: torch.Tensor) -> None:
        dec = self.decoder
        emb, acc = dec.embedding_dim, eng.acc_dtype
        fuse_w = dec.linear_fuse[0].weight
        sums = torch.empty(2 * emb, dtype=acc, device=dz.device)
        ops.bn_stats(dz, sums)  # column sums of dZ (no pivot): the gradient of the constant term sum_l W_l b_l
        colsum = sums[:emb]
        gwf = torch.empty((emb, 4 * emb), dtype=acc, device=dz.device) if fuse_w.requires_grad else None
        for slot, lv in enumerate(levels):
            f, lin = lv.feat, lv.lin
            cin = lv.m.shape[1]
            dy = dz if lv.lvl == 1 else ops.bilinear_bwd(dz, f.t.shape[1], f.t.shape[2])
            dm = torch.zeros((emb, cin), dtype=acc, device=dz.device)
            ops.conv2d_wgrad([f.t], dy, 1, 1, 0, 0, dm)
            if f.needs_grad:
                wt = ops.pack_conv_weight(lv.m.view(emb, cin, 1, 1), eng.dtype, 1)
                f.gsrcs.append((ops.conv2d_fwd([dy], wt, cin, 1, 1, 0, 0), 0))
            if lin.weight.requires_grad:
                eng.grad_buffer(lin.weight, False).copy_(torch.matmul(lv.wl.t(), dm))
            if lin.bias.requires_grad:
                eng.grad_buffer(lin.bias, False).copy_(torch.mv(lv.wl.t(), colsum))
            if gwf is not None:
                gwf[:, slot * emb:(slot + 1) * emb] = torch.addmm(torch.outer(colsum, lin.bias.detach().to(acc)), dm,
                                                                   lin.weight.detach().to(acc).t())
        if gwf is not None:
            eng.grad_buffer(fuse_w, False).view(emb, 4 * emb).copy_(gwf)

    # ====================================================================================== backward
    def _pgrads_ln(self, eng: Engine, ln: nn.LayerNorm):
        need = ln.weight.requires_grad or ln.bias.requires_grad
        return torch.zeros((2, ln.weight.numel()), dtype=eng.acc_dtype, device=ln.weight.device) if need else None

    def _store_ln_grads(self, eng: Engine, ln: nn.LayerNorm, pg) -> None:
        if pg is None:
            return
        if ln.weight.requires_grad:
            eng.grad_buffer(ln.weight, False).copy_(pg[0])
        if ln.bias.requires_grad:
            eng.grad_buffer(ln.bias, False).copy_(pg[1])

    @staticmethod
    def _take(act: Act) -> torch.Tensor:
        """the single registered gradient source of an activation (contiguous rows or a channel slice)"""
        assert len(act.gsrcs) == 1 and act.gsrcs[0][1] == 0
        g = act.gsrcs[0][0]
        act.gsrcs.clear()
        return g

    def _mlp_bwd(self, eng: Engine, blk: _Block, sv, g16: torch.Tensor, g32: torch.Tensor) -> torch.Tensor:
        """g16 / g32: stream gradient (16-bit copy / fp32).  Returns d(LN2 output) as 16-bit."""
        eng.conv_backward(sv.rc2, self._branch_grad(sv.rc2, sv.s, g16, g32))
        dy = self._take(sv.act_y)
        dw = blk.mlp.dwconv.dwconv
        c4 = dw.weight.shape[0]
        pg = torch.zeros((c4, 10), dtype=eng.acc_dtype, device=dy.device)
        df1 = ops.dwconv3x3_gelu_bwd(dy, sv.pre, sv.rc1.x, dw.weight.detach().view(c4, 9), pg)
        if dw.weight.requires_grad:
            eng.grad_buffer(dw.weight, False).view(c4, 9).copy_(pg[:, :9])
        if dw.bias is not None and dw.bias.requires_grad:
            eng.grad_buffer(dw.bias, False).copy_(pg[:, 9])
        eng.conv_backward(sv.rc1, df1)
        return self._take(sv.act_a)

    def _attention_bwd(self, eng: Engine, blk: _Block, sv, g16: torch.Tensor, g32: torch.Tensor) -> torch.Tensor:
        """g16 / g32: stream gradient (16-bit / fp32).  Returns d(LN1 output) as 16-bit (sum of the q and k/v paths)."""
        attn = blk.attn
        heads, d, nk, lp = sv.heads, sv.d, sv.nk, sv.lp
        eng.conv_backward(sv.rc_proj, self._branch_grad(sv.rc_proj, sv.s, g16, g32))
        do = self._take(sv.act_o)  # (B,h,w,C)
        b, h, w, c = do.shape
        n = h * w
        dt = eng.dtype
        do4 = do.view(b, 1, n, c)
        dkv32 = torch.zeros((b, lp, 2 * c), dtype=eng.acc_dtype, device=do.device)
        if ops.option("sra_fused") and ops.sra_attention_supported(n, nk, d, save_p=True):
            # ONE kernel: dP = dO.V^T in TMEM -> dS over the loaded P tile -> dQ = dS.K (csrc/sra_attention.cu); dP never reaches HBM
            dq3, ds3 = ops.sra_attention_bwd(do.view(b, n, c), sv.kv2, sv.p4.view(b, n, heads * lp), heads, nk, d ** -0.5)
            dq, ds4 = dq3.view(b, 1, n, c), ds3.view(b, 1, n, heads * lp)
        else:
            dp = torch.empty((b, 1, n, heads * lp), dtype=dt, device=do.device)
            ops.conv2d_fwd([do4[..., 0:d]], sv.kv2[:, c:c + d], nk, 1, 1, 0, 0, out=dp[..., 0:nk], w_rows_per_img=nk,
                           groups=(heads, d, d, lp))
            ds = ops.softmax_bwd(sv.p4.view(b, n, heads, lp), dp.view(b, n, heads, lp), d ** -0.5, nk)
            ds4 = ds.view(b, 1, n, heads * lp)
            dq = torch.empty((b, 1, n, c), dtype=dt, device=do.device)
            ops.conv2d_fwd([ds4[..., 0:lp]], sv.kv2[:, 0:d], d, 1, 1, 0, 0, out=dq[..., 0:d], w_rows_per_img=nk,
                           w_mn_major=True, groups=(heads, lp, d, d))
        # dV[b] = P^T dO,  dK[b] = dS^T q   (one independent product per image and head)
        if ops.option("attn_wgrad_grouped") and heads > 1:
            # all heads in one launch each: head g reads the dO / q columns shifted by g*d, the P / dS columns by g*lp
            ops.conv2d_wgrad([do4[..., 0:d]], sv.p4[..., 0:lp], 1, 1, 0, 0, dkv32[:, :, c:c + d], groups=(heads, d, lp, d))
            ops.conv2d_wgrad([sv.q4[..., 0:d]], ds4[..., 0:lp], 1, 1, 0, 0, dkv32[:, :, 0:d], groups=(heads, d, lp, d))
        else:
            for hd in range(heads):
                ops.conv2d_wgrad([do4[..., hd * d:(hd + 1) * d]], sv.p4[..., hd * lp:(hd + 1) * lp], 1, 1, 0, 0,
                                 dkv32[:, :, c + hd * d:c + (hd + 1) * d])
                ops.conv2d_wgrad([sv.q4[..., hd * d:(hd + 1) * d]], ds4[..., hd * lp:(hd + 1) * lp], 1, 1, 0, 0,
                                 dkv32[:, :, hd * d:(hd + 1) * d])
        dkv = ops.cast_f32(dkv32 if lp == nk else dkv32[:, :nk].contiguous(), dt)
        eng.conv_backward(sv.rc_kv, dkv.view(sv.rc_kv.x.shape))
        dkvin = self._take(sv.act_kvin)
        if attn.sr_ratio > 1:
            pg = self._pgrads_ln(eng, attn.norm)
            _, dsr = ops.layernorm_bwd(dkvin, sv.rc_sr.x, sv.st_sr, attn.norm.weight, want32=False, dtype16=dt, pgrads=pg)
            self._store_ln_grads(eng, attn.norm, pg)
            eng.conv_backward(sv.rc_sr, dsr)
            other = self._take(sv.act_a)  # gradient that reached LN1's output through sr -> kv
        else:
            other = dkvin
        eng.conv_backward(sv.rc_q, dq.view(b, h, w, c), dgrad_residual=other)
        return self._take(sv.act_a)

    def fused_train(self, eng: Engine, x16: torch.Tensor, bands: int, target: torch.Tensor, spec,
                    grad_scale: torch.Tensor | None = None) -> torch.Tensor:
        """FusedTrainer hook: normalised NHWC tiles -> loss, gradients left in the engine's destination buffers.  The final
        bilinear x4 (segformer.py:47-57) is fused with the loss: the (N,H,W,K) logits and their gradient never exist."""
        lr = self.run(eng, Act(x16, needs_grad=False), bands, upsample=not ops.option("fused_head"))
        n, h, w, k = lr.shape
        if not ops.option("fused_head"):
            coeff, _ = ops.seg_loss_fwd(lr, target, spec)
            d = torch.empty_like(lr)
            ops.seg_loss_bwd(lr, target, spec, coeff, grad_scale, d)
            self.backward(eng, d)
            return coeff[0]
        coeff, _ = ops.upsample_ce_fwd(lr, target, spec)
        d16 = torch.zeros((n, h, w, (k + 15) // 16 * 16), dtype=eng.dtype, device=lr.device)
        ops.upsample_ce_bwd(lr, target, spec, coeff, grad_scale, d16)
        self.backward(eng, None, d16_lowres=d16)
        return coeff[0]

    @torch.no_grad()
    def predict_classes(self, img: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
        """Eval post-processing of the reference's validation / test steps (segmentation_segformer.py:268-271,288-291):
        softmax(dim=1).argmax(dim=1) (K == 1: sigmoid > threshold) of forward(img), as ONE pass over the low-resolution
        logits — (N,H,W) int64 class map."""
        ops.require_cuda(img, "gdl_b200.SegFormer")
        eng = Engine(self.compute_dtype, training=False, wcache=self._wcache)
        lr = self.run(eng, self._input(img), img.shape[1], upsample=False)
        return ops.upsample_argmax(lr, img.shape[2], img.shape[3], threshold)

    def backward(self, eng: Engine, dlogits: torch.Tensor | None, d16_lowres: torch.Tensor | None = None) -> None:
        """dlogits: fp32 (N,H,W,K) gradient of the loss w.r.t. the logits returned by run(); or d16_lowres: the 16-bit,
        16-channel-padded gradient w.r.t. the low-resolution logits (fused head)."""
        S = eng.saved_segformer
        dt = eng.dtype
        if d16_lowres is not None:
            d16 = d16_lowres
        else:
            k = dlogits.shape[3]
            d128 = ops.bilinear_bwd(dlogits, S.h1, S.w1)
            d16 = ops.normalize_to_nhwc(d128, False, dt, (k + 15) // 16 * 16)
        eng.conv_backward(S.rc_pred, d16)
        eng.backward()  # linear_fuse BN/ReLU + the fuse GEMM: registers gradients on the 4 projections
        feat_grads: list[list[torch.Tensor]] = [[] for _ in range(4)]
        if S.projs is None:  # folded decoder: its backward (run from the tape) already registered d(feature) of every level
            for lvl in (4, 3, 2, 1):
                feat_grads[lvl - 1].append(self._take(S.stages[lvl - 1].feat))
        else:
            for p in S.projs:
                g = self._take(p.up)
                if p.lvl != 1:
                    g = ops.bilinear_bwd(g, p.rc.x.shape[1], p.rc.x.shape[2])
                eng.conv_backward(p.rc, g)
                feat_grads[p.lvl - 1].append(self._take(S.stages[p.lvl - 1].feat))
        for s in (3, 2, 1, 0):
            st = S.stages[s]
            gs = feat_grads[s]
            if len(gs) == 2:
                g = torch.empty(st.feat.t.shape, dtype=dt, device=st.feat.t.device)
                ops.grad_gather([(gs[0], 0), (gs[1], 0)], st.feat.t.shape, dt, g=g)
            else:
                g = gs[0]
            pg = self._pgrads_ln(eng, st.norm)
            gstream, g16 = ops.layernorm_bwd(g, st.x_out, st.st_out, st.norm.weight, want32=True, dtype16=dt, pgrads=pg)
            self._store_ln_grads(eng, st.norm, pg)
            for bs in reversed(st.blocks):
                blk = bs.blk
                da2 = self._mlp_bwd(eng, blk, bs.mlp, g16, gstream)
                pg = self._pgrads_ln(eng, blk.norm2)
                gstream, g16 = ops.layernorm_bwd(da2, bs.x_mid, bs.st2, blk.norm2.weight, add=gstream, want32=True,
                                                 dtype16=dt, pgrads=pg)
                self._store_ln_grads(eng, blk.norm2, pg)
                da1 = self._attention_bwd(eng, blk, bs.attn, g16, gstream)
                pg = self._pgrads_ln(eng, blk.norm1)
                gstream, g16 = ops.layernorm_bwd(da1, bs.x_in, bs.st1, blk.norm1.weight, add=gstream, want32=True,
                                                 dtype16=dt, pgrads=pg)
                self._store_ln_grads(eng, blk.norm1, pg)
            if st.dyn is not None:
                self._dyn_embed_bwd(eng, st.dyn, gstream)
                continue
            pg = self._pgrads_ln(eng, st.pe.norm)
            _, dpe = ops.layernorm_bwd(gstream, st.rc_pe.x, st.st_pe, st.pe.norm.weight, want32=False, dtype16=dt, pgrads=pg)
            self._store_ln_grads(eng, st.pe.norm, pg)
            eng.conv_backward(st.rc_pe, dpe)
            if s > 0:
                feat_grads[s - 1].append(self._take(S.stages[s - 1].feat))
        eng.saved_segformer = None

    # ====================================================================================== nn.Module surface
    def _input(self, image: torch.Tensor) -> Act:
        c = image.shape[1]
        x = ops.normalize_to_nhwc(image.contiguous().float(), True, self.compute_dtype, (c + 7) // 8 * 8)
        return Act(x, needs_grad=False)

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        ops.require_cuda(img, "gdl_b200.SegFormer")
        params = list(self.parameters())
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in params):
            return _SegFormerFn.apply(self, img, *params)
        with torch.no_grad():
            eng = Engine(self.compute_dtype, training=False, wcache=self._wcache)
            logits = self.run(eng, self._input(img), img.shape[1])
        return logits.permute(0, 3, 1, 2)


class _SegFormerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: SegFormer, image: torch.Tensor, *params: torch.Tensor) -> torch.Tensor:
        eng = Engine(model.compute_dtype, training=True, wcache=model._wcache, sync_bn_group=model.sync_bn_group)
        logits = model.run(eng, model._input(image), image.shape[1])
        ctx.eng, ctx.model, ctx.params = eng, model, params
        model.last_engine = eng
        return logits.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dlogits: torch.Tensor):
        eng: Engine = ctx.eng
        ctx.model.backward(eng, dlogits.permute(0, 2, 3, 1).contiguous().float())
        grads = tuple(eng.param_grads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        ctx.eng = None
        return (None, None, *grads)
