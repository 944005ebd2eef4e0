"""B200-native UNet++ (ResNet encoder + nested dense decoder + 3x3 head).

Drop-in for `smp.UnetPlusPlus(encoder_name, in_channels, encoder_weights, classes)` as built by
the reference at geo_deep_learning/tasks_with_models/segmentation_unetplus.py:126-131: same
constructor keywords, same `state_dict` keys/shapes (`encoder.conv1.weight`,
`decoder.blocks.x_0_0.conv1.0.weight`, `segmentation_head.0.weight`, ...), same forward
contract (float NCHW image in, (N, classes, H, W) logits out).  The nn.Conv2d / nn.BatchNorm2d
sub-modules are *parameter containers only*: their forward is never called.  All arithmetic
runs through gdl_b200.engine (tcgen05 implicit-GEMM convs, fused BN/ReLU/upsample kernels,
virtual concat) and the backward is the engine's hand-written one, exposed to PyTorch through
a single autograd.Function so `loss.backward()` / Lightning / DDP keep working.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..engine import Act, BNParams, Engine

_RESNETS = {
    # name: (block, layers, out_channels)
    "resnet18": ("basic", (2, 2, 2, 2), (64, 64, 128, 256, 512)),
    "resnet34": ("basic", (3, 4, 6, 3), (64, 64, 128, 256, 512)),
    "resnet50": ("bottleneck", (3, 4, 6, 3), (64, 256, 512, 1024, 2048)),
    "resnet101": ("bottleneck", (3, 4, 23, 3), (64, 256, 512, 1024, 2048)),
    "resnet152": ("bottleneck", (3, 8, 36, 3), (64, 256, 512, 1024, 2048)),
    # grouped 3x3 convolutions (torchvision ResNeXt: groups, width_per_group) — the encoder of the reference's shipped
    # UNet++ YAML (configs/unetplus_config_RGB.yaml:37 `encoder: resnext101_32x8d`)
    "resnext50_32x4d": ("bottleneck", (3, 4, 6, 3), (64, 256, 512, 1024, 2048), 32, 4),
    "resnext101_32x4d": ("bottleneck", (3, 4, 23, 3), (64, 256, 512, 1024, 2048), 32, 4),
    "resnext101_32x8d": ("bottleneck", (3, 4, 23, 3), (64, 256, 512, 1024, 2048), 32, 8),
}
DECODER_CHANNELS = (256, 128, 64, 32, 16)


def _conv(cin: int, cout: int, k: int, stride: int = 1, pad: int = 0) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False)


class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int) -> None:
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(_conv(inplanes, planes, 1, stride), nn.BatchNorm2d(planes))
        self.stride = stride


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int, groups: int = 1, base_width: int = 64) -> None:
        super().__init__()
        width = int(planes * (base_width / 64.0)) * groups  # torchvision.models.resnet.Bottleneck
        self.groups = groups
        self.conv1 = _conv(inplanes, width, 1)
        self.bn1 = nn.BatchNorm2d(width)
        # torchvision "v1.5": stride on the 3x3; ResNeXt: `groups` independent 3x3 convs over width / groups channels each
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = _conv(width, planes * 4, 1)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if stride != 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(_conv(inplanes, planes * 4, 1, stride), nn.BatchNorm2d(planes * 4))
        self.stride = stride


class ResNetEncoder(nn.Module):
    """Parameter layout of torchvision ResNet without avgpool/fc (what smp's ResNetEncoder keeps)."""

    def __init__(self, name: str, in_channels: int) -> None:
        super().__init__()
        if name not in _RESNETS:
            raise KeyError(f"Wrong encoder name `{name}`, supported encoders: {list(_RESNETS)}")
        kind, layers, self.out_channels = _RESNETS[name][:3]
        groups, wpg = (_RESNETS[name][3:] + (1, 64))[:2] if len(_RESNETS[name]) > 3 else (1, 64)
        self.in_channels = in_channels
        if kind == "basic":
            block = _BasicBlock
        else:
            def block(inplanes: int, planes: int, stride: int) -> _Bottleneck:
                return _Bottleneck(inplanes, planes, stride, groups, wpg)
            block.expansion = _Bottleneck.expansion
        self.conv1 = _conv(in_channels, 64, 7, 2, 3)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for i, (planes, nblk) in enumerate(zip((64, 128, 256, 512), layers)):
            blocks = []
            for b in range(nblk):
                blocks.append(block(inplanes, planes, (1 if i == 0 else 2) if b == 0 else 1))
                inplanes = planes * block.expansion
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        for m in self.modules():  # torchvision.models.resnet.ResNet.__init__
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def _conv_bn(cin: int, cout: int) -> nn.Sequential:
    # smp Conv2dReLU = Sequential(conv, bn, relu): keep index 0 / 1 for the keys
    return nn.Sequential(_conv(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecoderBlock(nn.Module):
    def __init__(self, in_ch: int, skip_ch: int, out_ch: int) -> None:
        super().__init__()
        self.conv1 = _conv_bn(in_ch + skip_ch, out_ch)
        self.conv2 = _conv_bn(out_ch, out_ch)


def decoder_plan(enc_ch: tuple[int, ...]) -> dict[str, tuple[int, int, int]]:
    enc = list(enc_ch)[::-1]
    in_ch = [enc[0]] + list(DECODER_CHANNELS[:-1])
    skip_ch = enc[1:] + [0]
    out_ch = list(DECODER_CHANNELS)
    plan: dict[str, tuple[int, int, int]] = {}
    for layer in range(len(in_ch) - 1):
        for depth in range(layer + 1):
            if depth == 0:
                plan[f"x_{depth}_{layer}"] = (in_ch[layer], skip_ch[layer] * (layer + 1), out_ch[layer])
            else:
                plan[f"x_{depth}_{layer}"] = (skip_ch[layer - 1], skip_ch[layer] * (layer + 1 - depth), skip_ch[layer])
    plan[f"x_0_{len(in_ch) - 1}"] = (in_ch[-1], 0, out_ch[-1])
    return plan


class _Decoder(nn.Module):
    def __init__(self, enc_ch: tuple[int, ...]) -> None:
        super().__init__()
        self.blocks = nn.ModuleDict({k: _DecoderBlock(*v) for k, v in decoder_plan(enc_ch).items()})
        for m in self.modules():  # smp.base.initialization.initialize_decoder
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def _bnp(bn: nn.BatchNorm2d) -> BNParams:
    return BNParams(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.eps,
                    bn.momentum if bn.momentum is not None else 0.1)


class UnetPlusPlus(nn.Module):
    def __init__(self, encoder_name: str = "resnet34", in_channels: int = 3, encoder_weights: str | None = None,
                 classes: int = 1, compute_dtype: torch.dtype = torch.bfloat16, **kwargs: object) -> None:
        super().__init__()
        if encoder_weights is not None:
            # the reference downloads ImageNet weights here; there is no network in this build.
            # Load them afterwards with load_state_dict (keys are identical).
            raise ValueError("encoder_weights must be None: load pretrained tensors through load_state_dict")
        self.encoder = ResNetEncoder(encoder_name, in_channels)
        self.decoder = _Decoder(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(DECODER_CHANNELS[-1], classes, 3, padding=1))
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)
        nn.init.constant_(self.segmentation_head[0].bias, 0)
        self.classes = classes
        self.compute_dtype = compute_dtype
        self.last_engine: Engine | None = None
        self._wcache: dict = {}
        self.sync_bn_group = None  # set to a torch.distributed group for SyncBatchNorm statistics

    # ------------------------------------------------------------------ engine graph
    def _block(self, eng: Engine, blk: nn.Module, x: Act, want_up: bool) -> Act:
        if isinstance(blk, _BasicBlock):
            b1, b2 = _bnp(blk.bn1), _bnp(blk.bn2)
            r1 = eng.conv_raw([x], blk.conv1.weight, blk.stride, 1, stats_for=b1)
            a1 = eng.bn_act(r1, eng.bn_prepare(r1, b1))
            r2 = eng.conv_raw([a1], blk.conv2.weight, 1, 1, stats_for=b2)
            last, last_bn = r2, b2
        else:
            b1, b2, b3 = _bnp(blk.bn1), _bnp(blk.bn2), _bnp(blk.bn3)
            r1 = eng.conv_raw([x], blk.conv1.weight, 1, 0, stats_for=b1)
            a1 = eng.bn_act(r1, eng.bn_prepare(r1, b1))
            r2 = (eng.conv_raw([a1], blk.conv2.weight, blk.stride, 1, stats_for=b2) if blk.groups == 1
                  else eng.conv_raw_grouped(a1, blk.conv2.weight, blk.groups, blk.stride, 1))
            a2 = eng.bn_act(r2, eng.bn_prepare(r2, b2))
            last = eng.conv_raw([a2], blk.conv3.weight, 1, 0, stats_for=b3)
            last_bn = b3
        bn_last = eng.bn_prepare(last, last_bn)
        if blk.downsample is not None:
            bd = _bnp(blk.downsample[1])
            rd = eng.conv_raw([x], blk.downsample[0].weight, blk.stride, 0, stats_for=bd)
            bnd = eng.bn_prepare(rd, bd)
            return eng.bn_act(last, bn_last, res_branch=(rd, bnd), want_up=want_up)
        return eng.bn_act(last, bn_last, residual=x, want_up=want_up)

    def _layer(self, eng: Engine, layer: nn.Sequential, x: Act, want_up: bool) -> Act:
        n = len(layer)
        for i, blk in enumerate(layer):
            x = self._block(eng, blk, x, want_up and i == n - 1)
        return x

    def _decoder_block(self, eng: Engine, name: str, x: Act, skips: list[Act], want_up: bool) -> Act:
        blk = self.decoder.blocks[name]
        if x.up is None:
            raise RuntimeError(f"{name}: producer did not emit an upsampled copy")
        b1, b2 = _bnp(blk.conv1[1]), _bnp(blk.conv2[1])
        r1 = eng.conv_raw([x.up, *skips], blk.conv1[0].weight, 1, 1, stats_for=b1)
        a1 = eng.bn_act(r1, eng.bn_prepare(r1, b1))
        r2 = eng.conv_raw([a1], blk.conv2[0].weight, 1, 1, stats_for=b2)
        return eng.bn_act(r2, eng.bn_prepare(r2, b2), want_up=want_up)

    def run(self, eng: Engine, x: Act) -> torch.Tensor:
        """x: NHWC 16-bit input (channels possibly zero-padded). Returns fp32 logits (N,H,W,K)."""
        enc = self.encoder
        _, h, w, _ = x.t.shape
        if h % 32 or w % 32:
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width divisible by 32.")
        bn1 = _bnp(enc.bn1)
        r = eng.conv_raw([x], enc.conv1.weight, 2, 3, stats_for=bn1)
        e1 = eng.bn_act(r, eng.bn_prepare(r, bn1))  # 64 @ H/2
        p = eng.maxpool3x3s2(e1)
        e2 = self._layer(eng, enc.layer1, p, True)
        e3 = self._layer(eng, enc.layer2, e2, True)
        e4 = self._layer(eng, enc.layer3, e3, True)
        e5 = self._layer(eng, enc.layer4, e4, True)
        feats = [e5, e4, e3, e2, e1]
        dense: dict[str, Act] = {}
        n = len(feats) - 1
        for layer in range(n):
            for depth in range(n - layer):
                dl = depth + layer
                # x_{depth}_{dl} is upsampled by x_{depth}_{dl+1} (exists while dl < n-1) or by x_0_n
                up = dl < n - 1 or depth == 0
                if layer == 0:
                    dense[f"x_{depth}_{depth}"] = self._decoder_block(eng, f"x_{depth}_{depth}", feats[depth],
                                                                      [feats[depth + 1]], up)
                else:
                    skips = [dense[f"x_{i}_{dl}"] for i in range(depth + 1, dl + 1)] + [feats[dl + 1]]
                    dense[f"x_{depth}_{dl}"] = self._decoder_block(eng, f"x_{depth}_{dl}", dense[f"x_{depth}_{dl - 1}"],
                                                                   skips, up)
        last = self._decoder_block(eng, f"x_0_{n}", dense[f"x_0_{n - 1}"], [], False)
        dense[f"x_0_{n}"] = last
        eng.named = {"e1": e1, "e2": e2, "e3": e3, "e4": e4, "e5": e5, **dense}  # for tests / debugging
        head = self.segmentation_head[0]
        return eng.conv_head(last, head.weight, head.bias, 1)

    # ------------------------------------------------------------------ nn.Module surface
    def _input(self, image: torch.Tensor) -> Act:
        c = image.shape[1]
        ld = (c + 7) // 8 * 8
        x = ops.normalize_to_nhwc(image.contiguous().float(), True, self.compute_dtype, ld)
        return Act(x, needs_grad=False)

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        ops.require_cuda(image, "gdl_b200.UnetPlusPlus")
        params = [p for p in self.parameters()]
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in params):
            return _UnetPPFn.apply(self, image, *params)
        with torch.no_grad():
            eng = Engine(self.compute_dtype, training=False, wcache=self._wcache)
            logits = self.run(eng, self._input(image))
        return logits.permute(0, 3, 1, 2)


class _UnetPPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: UnetPlusPlus, image: torch.Tensor, *params: torch.Tensor) -> torch.Tensor:
        eng = Engine(model.compute_dtype, training=True, wcache=model._wcache, sync_bn_group=model.sync_bn_group)
        logits = model.run(eng, model._input(image))
        ctx.eng = eng
        ctx.params = params
        model.last_engine = eng
        return logits.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dlogits: torch.Tensor):
        eng: Engine = ctx.eng
        d = dlogits.permute(0, 2, 3, 1).contiguous().float()
        k = d.shape[3]
        d16 = ops.normalize_to_nhwc(d, False, eng.dtype, (k + 15) // 16 * 16)
        eng.head_backward(d16)
        eng.backward()
        grads = tuple(eng.param_grads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        ctx.eng = None
        return (None, None, *grads)
