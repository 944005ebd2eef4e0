"""ORACLE (test infrastructure — never imported by the product path).

CPU/fp32 restatement of `segmentation_models_pytorch==0.5.0` `UnetPlusPlus`, the model the
reference builds at geo_deep_learning/tasks_with_models/segmentation_unetplus.py:126-131
(`smp.UnetPlusPlus(encoder_name, in_channels, encoder_weights, classes)`).  smp is an
un-vendored, pinned dependency (uv.lock.cu128:3250) that is absent from /root/reference and
cannot be installed here (no network), so its published algorithm is restated from the
structure recorded in SURVEY.md Appendix B:

  * encoder  = torchvision ResNet without avgpool/fc; features
               [x, relu(bn1(conv1 x)), layer1(maxpool .), layer2, layer3, layer4]
  * decoder  = 11 nested blocks x_{depth}_{layer}; each block: nearest x2 upsample ->
               cat(skip) -> (Conv3x3 no-bias + BN + ReLU) x 2
  * head     = Conv3x3(16 -> classes, bias)

PARITY STATUS: **unpinned for values** — the reference's only test at this boundary
(tests/test_notebooks_00quickstart.py:55-71,101-118) asserts no values.  Shapes are pinned:
resnet34 / 3 bands / 2 classes has 26 078 754 parameters == the notebook's recorded "26.1 M"
(notebooks/00_quickstart.ipynb:572-576); see tests/test_oracle_cpu.py.  The encoder half IS
pinned because it executes torchvision's own ResNet modules.

state_dict keys follow smp (`encoder.conv1.weight`, `decoder.blocks.x_0_0.conv1.0.weight`,
`segmentation_head.0.weight`, ...) so checkpoints are interchangeable with the product model.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

_ENCODERS = {
    # name: (torchvision ctor, out_channels after the input)
    "resnet18": (torchvision.models.resnet18, (64, 64, 128, 256, 512)),
    "resnet34": (torchvision.models.resnet34, (64, 64, 128, 256, 512)),
    "resnet50": (torchvision.models.resnet50, (64, 256, 512, 1024, 2048)),
    "resnet101": (torchvision.models.resnet101, (64, 256, 512, 1024, 2048)),
    "resnext50_32x4d": (torchvision.models.resnext50_32x4d, (64, 256, 512, 1024, 2048)),
    "resnext101_32x8d": (torchvision.models.resnext101_32x8d, (64, 256, 512, 1024, 2048)),
}
DECODER_CHANNELS = (256, 128, 64, 32, 16)


def encoder_channels(name: str) -> tuple[int, ...]:
    return _ENCODERS[name][1]


def make_encoder(name: str, in_channels: int) -> nn.Module:
    """torchvision ResNet minus avgpool/fc; first conv re-created for in_channels != 3
    (same rule as geo_deep_learning/models/utils.py:140-181 with pretrained=False)."""
    net = _ENCODERS[name][0](weights=None)
    del net.fc
    del net.avgpool
    if in_channels != 3:
        old = net.conv1
        net.conv1 = nn.Conv2d(in_channels, old.out_channels, kernel_size=7, stride=2, padding=3, bias=False)
    return net


def encoder_features(enc: nn.Module, x: torch.Tensor) -> list[torch.Tensor]:
    f0 = x
    f1 = enc.relu(enc.bn1(enc.conv1(x)))
    f2 = enc.layer1(enc.maxpool(f1))
    f3 = enc.layer2(f2)
    f4 = enc.layer3(f3)
    f5 = enc.layer4(f4)
    return [f0, f1, f2, f3, f4, f5]


def _conv_bn_relu(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class DecoderBlock(nn.Module):
    def __init__(self, in_ch: int, skip_ch: int, out_ch: int) -> None:
        super().__init__()
        self.conv1 = _conv_bn_relu(in_ch + skip_ch, out_ch)
        self.conv2 = _conv_bn_relu(out_ch, out_ch)

    def forward(self, x: torch.Tensor, skip: torch.Tensor | None = None) -> torch.Tensor:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


def decoder_plan(enc_ch: tuple[int, ...]) -> dict[str, tuple[int, int, int]]:
    """name -> (in_ch, skip_ch, out_ch) for the 11 blocks (smp 0.5.0 UnetPlusPlusDecoder.__init__)."""
    enc = list(enc_ch)[::-1]  # deepest first: e5,e4,e3,e2,e1
    in_ch = [enc[0]] + list(DECODER_CHANNELS[:-1])
    skip_ch = enc[1:] + [0]
    out_ch = list(DECODER_CHANNELS)
    plan = {}
    for layer in range(len(in_ch) - 1):
        for depth in range(layer + 1):
            if depth == 0:
                i, s, o = in_ch[layer], skip_ch[layer] * (layer + 1), out_ch[layer]
            else:
                o = skip_ch[layer]
                s = skip_ch[layer] * (layer + 1 - depth)
                i = skip_ch[layer - 1]
            plan[f"x_{depth}_{layer}"] = (i, s, o)
    plan[f"x_0_{len(in_ch) - 1}"] = (in_ch[-1], 0, out_ch[-1])
    return plan


class Decoder(nn.Module):
    def __init__(self, enc_ch: tuple[int, ...]) -> None:
        super().__init__()
        self.depth = len(enc_ch) - 1  # 4
        self.blocks = nn.ModuleDict({k: DecoderBlock(*v) for k, v in decoder_plan(enc_ch).items()})

    def forward(self, feats: list[torch.Tensor]) -> torch.Tensor:
        feats = feats[1:][::-1]  # drop the input, deepest first
        dense: dict[str, torch.Tensor] = {}
        n = len(feats) - 1  # 4
        for layer in range(n):
            for depth in range(n - layer):
                if layer == 0:
                    dense[f"x_{depth}_{depth}"] = self.blocks[f"x_{depth}_{depth}"](feats[depth], feats[depth + 1])
                else:
                    dl = depth + layer
                    cat = [dense[f"x_{i}_{dl}"] for i in range(depth + 1, dl + 1)]
                    cat = torch.cat(cat + [feats[dl + 1]], dim=1)
                    dense[f"x_{depth}_{dl}"] = self.blocks[f"x_{depth}_{dl}"](dense[f"x_{depth}_{dl - 1}"], cat)
        dense[f"x_0_{n}"] = self.blocks[f"x_0_{n}"](dense[f"x_0_{n - 1}"])
        return dense[f"x_0_{n}"]


class UnetPlusPlusOracle(nn.Module):
    def __init__(self, encoder_name: str = "resnet50", in_channels: int = 3, classes: int = 1) -> None:
        super().__init__()
        self.encoder = make_encoder(encoder_name, in_channels)
        self.decoder = Decoder(encoder_channels(encoder_name))
        self.segmentation_head = nn.Sequential(nn.Conv2d(DECODER_CHANNELS[-1], classes, 3, padding=1))
        # smp.base.initialization: decoder convs kaiming-uniform(fan_in, relu), BN (1,0); head xavier-uniform
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)
        nn.init.constant_(self.segmentation_head[0].bias, 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[-1] % 32 or x.shape[-2] % 32:
            raise RuntimeError(f"Wrong input shape height={x.shape[-2]}, width={x.shape[-1]}: must be divisible by 32")
        return self.segmentation_head(self.decoder(encoder_features(self.encoder, x)))
