"""ORACLE (test infrastructure — never imported by the product path).

Functional CPU/fp32 restatement of the trainable half of the reference's DOFA configuration
(`freeze_layers: ["encoder"]`, configs/dofa_config_RGB.yaml:57): everything downstream of the 4 encoder
feature maps in `DOFASegmentationModel.forward` (geo_deep_learning/models/segmentation/dofa.py:83-107):

  neck      MultiLevelNeck            models/necks/multilevel_neck.py:139-160 (1x1 conv+bias+BN+ReLU,
                                      bilinear resize x[4,2,1,0.5], 3x3 conv+bias+BN+ReLU)
  decoder   UperNetDecoder            models/decoders/upernet.py:111-152 (PPM models/utils.py:55-93,
                                      laterals, top-down add, fpn convs, concat, fpn_bottleneck)
  head      SegmentationHead 1x1      models/heads/segmentation_head.py:19-26
  aux_head  FCNHead                   models/heads/fcn_head.py:75-84 (Dropout2d = identity here)
  both logits are bilinearly resized to the image size (align_corners=False).

Written over a plain state_dict with the reference's keys (`neck.lateral_convs.0.conv.weight`, ...).
PARITY STATUS: **pinned** — compared with the reference's own modules (all four are torch-only and
importable from /root/reference) in tests/test_oracle_cpu.py when the tree is present, and with
tests/golden/upernet_golden.pt (outputs of those modules, oracle/make_golden.py) where it is not.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SCALES = (4, 2, 1, 0.5)
POOL_SCALES = (1, 2, 3, 6)


def _cbr(sd, p, x, training, pad, momentum=0.1):
    """ConvModule: conv (bias optional) + BatchNorm2d + ReLU"""
    y = F.conv2d(x, sd[p + ".conv.weight"], sd.get(p + ".conv.bias"), padding=pad)
    y = F.batch_norm(y, sd[p + ".norm.running_mean"], sd[p + ".norm.running_var"], sd[p + ".norm.weight"],
                     sd[p + ".norm.bias"], training, momentum, 1e-5)
    return F.relu(y)


def _up(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def neck(sd, feats, training):
    outs = []
    for i, f in enumerate(feats):
        y = _cbr(sd, f"neck.lateral_convs.{i}", f, training, 0)
        h, w = y.shape[2:]
        y = _up(y, (int(h * SCALES[i]), int(w * SCALES[i])))  # resize(): size = int(h * scale) (models/utils.py:106-111)
        outs.append(_cbr(sd, f"neck.convs.{i}", y, training, 1))
    return outs


def decoder(sd, feats, training):
    x = feats[-1]
    psp = [x]
    for j, s in enumerate(POOL_SCALES):
        p = F.adaptive_avg_pool2d(x, s)
        p = _cbr(sd, f"decoder.psp_modules.{j}.1", p, training, 0)
        psp.append(_up(p, x.shape[2:]))
    lat = [_cbr(sd, f"decoder.lateral_convs.{i}", feats[i], training, 0) for i in range(3)]
    lat.append(_cbr(sd, "decoder.bottleneck", torch.cat(psp, 1), training, 1))
    for i in range(3, 0, -1):
        lat[i - 1] = lat[i - 1] + _up(lat[i], lat[i - 1].shape[2:])
    outs = [_cbr(sd, f"decoder.fpn_convs.{i}", lat[i], training, 1) for i in range(3)] + [lat[3]]
    size0 = outs[0].shape[2:]
    outs = [outs[0]] + [_up(o, size0) for o in outs[1:]]
    return _cbr(sd, "decoder.fpn_bottleneck", torch.cat(outs, 1), training, 1)


def upernet_forward(sd, enc_feats, image_size, training=False):
    """enc_feats: 4 x (B, C, h, w) encoder maps.  Returns (out, aux) logits at image_size."""
    feats = neck(sd, enc_feats, training)
    x = decoder(sd, feats, training)
    out = _up(F.conv2d(x, sd["head.conv.weight"], sd["head.conv.bias"]), image_size)
    a = _cbr(sd, "aux_head.convs.0", feats[-1], training, 1)
    aux = _up(F.conv2d(a, sd["aux_head.cls_seg.weight"], sd["aux_head.cls_seg.bias"]), image_size)
    return out, aux


def init_state_dict(embed_dim=768, channels=256, num_classes=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def cbr(p, cin, cout, k, bias):
        sd[p + ".conv.weight"] = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
        if bias:
            sd[p + ".conv.bias"] = torch.randn(cout, generator=g) * 0.05
        sd[p + ".norm.weight"] = 1 + 0.2 * torch.randn(cout, generator=g)
        sd[p + ".norm.bias"] = 0.1 * torch.randn(cout, generator=g)
        sd[p + ".norm.running_mean"] = torch.zeros(cout)
        sd[p + ".norm.running_var"] = torch.ones(cout)
        sd[p + ".norm.num_batches_tracked"] = torch.tensor(0)

    for i in range(4):
        cbr(f"neck.lateral_convs.{i}", embed_dim, embed_dim, 1, True)
    for i in range(4):
        cbr(f"neck.convs.{i}", embed_dim, embed_dim, 3, True)
    for j in range(4):
        cbr(f"decoder.psp_modules.{j}.1", embed_dim, channels, 1, False)
    cbr("decoder.bottleneck", embed_dim + 4 * channels, channels, 3, False)
    for i in range(3):
        cbr(f"decoder.lateral_convs.{i}", embed_dim, channels, 1, False)
    for i in range(3):
        cbr(f"decoder.fpn_convs.{i}", channels, channels, 3, False)
    cbr("decoder.fpn_bottleneck", 4 * channels, channels, 3, False)
    cbr("aux_head.convs.0", embed_dim, channels, 3, False)
    sd["aux_head.cls_seg.weight"] = torch.randn(num_classes, channels, 1, 1, generator=g) * 0.05
    sd["aux_head.cls_seg.bias"] = torch.randn(num_classes, generator=g) * 0.05
    sd["head.conv.weight"] = torch.randn(num_classes, channels, 1, 1, generator=g) * 0.05
    sd["head.conv.bias"] = torch.randn(num_classes, generator=g) * 0.05
    return sd
