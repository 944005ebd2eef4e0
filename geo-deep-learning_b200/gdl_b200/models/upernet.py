"""B200-native MultiLevelNeck + UperNetDecoder + SegmentationHead + FCNHead: the trainable half of the
reference's DOFA configuration (`freeze_layers: ["encoder"]`, configs/dofa_config_RGB.yaml:57), i.e.
everything `DOFASegmentationModel.forward` does after the encoder (models/segmentation/dofa.py:83-107).

State_dict keys equal the reference's (`neck.lateral_convs.0.conv.weight`, `decoder.psp_modules.0.1.conv.weight`,
`decoder.fpn_bottleneck.norm.running_mean`, `aux_head.cls_seg.bias`, `head.conv.weight`, ...).  The DOFA ViT
encoder itself (dynamic wavelength-conditioned patch embedding + 12 ViT blocks) is not built yet: this module
takes the 4 encoder feature maps as input (a14-a15 are next).

torch.cat of the PPM outputs (1792 channels) and of the 4 FPN levels (1024 channels) is a virtual concat
read in place by the tcgen05 conv; BatchNorm / ReLU / bilinear / adaptive-pool / add are the fused
HBM-bound kernels with hand-written backward (engine tape).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..engine import Act, Engine

SCALES = (4, 2, 1, 0.5)
POOL_SCALES = (1, 2, 3, 6)


class _ConvModule(nn.Module):
    """parameter container: conv (+bias) -> BatchNorm2d -> ReLU"""

    def __init__(self, cin: int, cout: int, k: int, bias: bool) -> None:
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=bias)
        self.norm = nn.BatchNorm2d(cout)


class MultiLevelNeck(nn.Module):
    def __init__(self, in_channels: list[int], out_channels: list[int], scales: list[float] | None = None) -> None:
        super().__init__()
        if not isinstance(in_channels, list):
            raise TypeError(f"in_channels must be a list, but got {type(in_channels)}")
        if not isinstance(out_channels, list):
            raise TypeError(f"out_channels must be a list, but got {type(out_channels)}")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.scales = list(scales or [0.5, 1, 2, 4])
        self.lateral_convs = nn.ModuleList(_ConvModule(i, o, 1, True) for i, o in zip(in_channels, out_channels))
        self.convs = nn.ModuleList(_ConvModule(o, o, 3, True) for o in out_channels)
        for m in self.modules():  # init_weights: xavier-uniform convs
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)


class UperNetDecoder(nn.Module):
    def __init__(self, embed_dim: list[int], pool_scales=POOL_SCALES, channels: int = 256) -> None:
        super().__init__()
        self.pool_scales = tuple(pool_scales)
        self.psp_modules = nn.ModuleList(nn.Sequential(nn.Identity(), _ConvModule(embed_dim[-1], channels, 1, False))
                                         for _ in self.pool_scales)
        self.bottleneck = _ConvModule(embed_dim[-1] + len(self.pool_scales) * channels, channels, 3, False)
        self.lateral_convs = nn.ModuleList(_ConvModule(e, channels, 1, False) for e in embed_dim[:-1])
        self.fpn_convs = nn.ModuleList(_ConvModule(channels, channels, 3, False) for _ in embed_dim[:-1])
        self.fpn_bottleneck = _ConvModule(len(embed_dim) * channels, channels, 3, False)


class FCNHead(nn.Module):
    def __init__(self, in_channels: int, channels: int, num_classes: int) -> None:
        super().__init__()
        self.convs = nn.Sequential(_ConvModule(in_channels, channels, 3, False))
        self.cls_seg = nn.Conv2d(channels, num_classes, 1)


class SegmentationHead(nn.Module):
    def __init__(self, in_channels: int, num_classes: int) -> None:
        super().__init__()
        self.conv = nn.Conv2d(in_channels, num_classes, 1)


class UperNetSegmentor(nn.Module):
    """neck + decoder + head + aux_head.  forward(enc_feats, image_size) -> (out, aux) logits (N,K,H,W)."""

    def __init__(self, embed_dim: int = 768, channels: int = 256, num_classes: int = 1,
                 compute_dtype: torch.dtype = torch.bfloat16, aux_dropout_ratio: float = 0.0) -> None:
        super().__init__()
        # FCNHead's nn.Dropout2d before cls_seg (fcn_head.py:69-83; the reference default is 0.1, active in train mode).
        # 0 keeps the deterministic behaviour parity is defined on; the task mirror passes the reference's value.
        self.aux_dropout_ratio = aux_dropout_ratio
        self.aux_dropout_mask: torch.Tensor | None = None  # (N, channels) override of the draw (tests)
        self.neck = MultiLevelNeck([embed_dim] * 4, [embed_dim] * 4, scales=list(SCALES))
        self.decoder = UperNetDecoder([embed_dim] * 4, POOL_SCALES, channels)
        self.aux_head = FCNHead(embed_dim, channels, num_classes)
        self.head = SegmentationHead(channels, num_classes)
        self.compute_dtype = compute_dtype
        self.num_classes = num_classes
        self._wcache: dict = {}
        self.sync_bn_group = None

    def run(self, eng: Engine, feats: list[Act], image_size: tuple[int, int], upsample: bool = True):
        """feats: 4 NHWC 16-bit encoder maps. Returns fp32 logits (out, aux), both (N, H, W, K); with upsample=False the two
        heads' own low-resolution maps (for the fused upsample + loss / argmax head)."""
        if len(feats) != len(self.neck.in_channels):
            raise ValueError(f"len(inputs) must be equal to len(in_channels), but got {len(feats)} and {len(self.neck.in_channels)}")
        cb = eng.conv_bn_relu
        # ---- MultiLevelNeck (multilevel_neck.py:139-160)
        nf = []
        for i, f in enumerate(feats):
            y = cb([f], self.neck.lateral_convs[i].conv, self.neck.lateral_convs[i].norm)
            h, w = y.t.shape[1:3]
            y = eng.bilinear(y, int(h * self.neck.scales[i]), int(w * self.neck.scales[i]))
            nf.append(cb([y], self.neck.convs[i].conv, self.neck.convs[i].norm))
        # ---- UperNetDecoder (upernet.py:111-152)
        dec = self.decoder
        x = nf[-1]
        hx, wx = x.t.shape[1:3]
        psp = [x]
        for j, s in enumerate(dec.pool_scales):
            pm = dec.psp_modules[j][1]
            psp.append(eng.bilinear(cb([eng.adaptive_avgpool(x, s)], pm.conv, pm.norm), hx, wx))
        lat = [cb([nf[i]], dec.lateral_convs[i].conv, dec.lateral_convs[i].norm) for i in range(3)]
        lat.append(cb(psp, dec.bottleneck.conv, dec.bottleneck.norm))  # virtual concat of 768 + 4 x 256 channels
        for i in range(3, 0, -1):
            lat[i - 1] = eng.add(lat[i - 1], eng.bilinear(lat[i], *lat[i - 1].t.shape[1:3]))
        outs = [cb([lat[i]], dec.fpn_convs[i].conv, dec.fpn_convs[i].norm) for i in range(3)] + [lat[3]]
        h0, w0 = outs[0].t.shape[1:3]
        outs = [outs[0]] + [eng.bilinear(o, h0, w0) for o in outs[1:]]
        y = cb(outs, dec.fpn_bottleneck.conv, dec.fpn_bottleneck.norm)  # virtual concat of 4 x 256 channels
        # ---- heads
        acc = eng.acc_dtype
        rc_out = eng.conv_raw([y], self.head.conv.weight, 1, 0, bias=self.head.conv.bias, out_dtype=acc)
        a = cb([nf[-1]], self.aux_head.convs[0].conv, self.aux_head.convs[0].norm)
        a = eng.dropout2d(a, self.aux_dropout_ratio, self.aux_dropout_mask)
        rc_aux = eng.conv_raw([a], self.aux_head.cls_seg.weight, 1, 0, bias=self.aux_head.cls_seg.bias, out_dtype=acc)
        eng.saved_upernet = (rc_out, rc_aux)  # per forward (engine), not per module
        eng.named = {f"neck{i}": nf[i] for i in range(4)} | {"fpn": y}
        if not upsample:
            return rc_out.x, rc_aux.x
        return ops.bilinear_fwd(rc_out.x, *image_size), ops.bilinear_fwd(rc_aux.x, *image_size)

    def backward(self, eng: Engine, d_out: torch.Tensor | None, d_aux: torch.Tensor | None, lowres16: bool = False) -> None:
        """d_out / d_aux: fp32 (N,H,W,K) gradients of the loss w.r.t. the two logit maps; with lowres16 they are the 16-bit,
        16-channel-padded gradients w.r.t. the heads' low-resolution maps (fused head)."""
        rc_out, rc_aux = eng.saved_upernet
        for rc, d in ((rc_out, d_out), (rc_aux, d_aux)):
            if d is None:
                continue
            if lowres16:
                eng.conv_backward(rc, d)
                continue
            k = d.shape[3]
            dl = ops.bilinear_bwd(d, rc.x.shape[1], rc.x.shape[2])
            eng.conv_backward(rc, ops.normalize_to_nhwc(dl, False, eng.dtype, (k + 15) // 16 * 16))
        eng.backward()
        eng.saved_upernet = None

    def forward(self, enc_feats: list[torch.Tensor], image_size: tuple[int, int]):
        """enc_feats: 4 x (N, C, h, w) float tensors (what DOFAv2.forward returns)."""
        ops.require_cuda(enc_feats[0], "gdl_b200.UperNetSegmentor")
        params = list(self.parameters())
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in params):
            out = _UperNetFn.apply(self, tuple(image_size), len(enc_feats), *enc_feats, *params)
            return out[0], out[1]
        with torch.no_grad():
            eng = Engine(self.compute_dtype, training=False, wcache=self._wcache)
            o, a = self.run(eng, [self._feat(f, False) for f in enc_feats], image_size)
        return o.permute(0, 3, 1, 2), a.permute(0, 3, 1, 2)

    def _feat(self, f: torch.Tensor, needs_grad: bool) -> Act:
        c = f.shape[1]
        return Act(ops.normalize_to_nhwc(f.contiguous().float(), True, self.compute_dtype, c), needs_grad=needs_grad)


class _UperNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: UperNetSegmentor, image_size, nfeat: int, *args: torch.Tensor):
        feats, params = args[:nfeat], args[nfeat:]
        eng = Engine(model.compute_dtype, training=True, wcache=model._wcache, sync_bn_group=model.sync_bn_group,
                     acc_dtype=getattr(model, "acc_dtype", torch.float32))
        acts = [model._feat(f, f.requires_grad) for f in feats]
        o, a = model.run(eng, acts, image_size)
        ctx.eng, ctx.model, ctx.params, ctx.acts = eng, model, params, acts
        return o.permute(0, 3, 1, 2), a.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, d_out: torch.Tensor, d_aux: torch.Tensor):
        eng: Engine = ctx.eng

        def nhwc(d):
            return None if d is None else d.permute(0, 2, 3, 1).contiguous().to(eng.acc_dtype)
        ctx.model.backward(eng, nhwc(d_out), nhwc(d_aux))
        fgrads = []
        for a in ctx.acts:  # gradients w.r.t. the encoder maps (None when the encoder is frozen)
            g = eng.collect_grad(a) if a.needs_grad else None
            fgrads.append(None if g is None else g.to(eng.acc_dtype).permute(0, 3, 1, 2))
        grads = tuple(eng.param_grads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        ctx.eng = None
        return (None, None, None, *fgrads, *grads)
