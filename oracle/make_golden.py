"""Generate tests/golden/*.pt (run in the BUILD container, where /root/reference exists).

  python oracle/make_golden.py

* tensors_golden.pt   — outputs of the REFERENCE's own geo_deep_learning/utils/tensors.py
                        (`normalization`, `standardization`) on seeded inputs: pins oracle/tensors.py
                        and the CUDA normalise kernel to the reference itself.
* wds_golden.pt       — outputs of the REFERENCE's `ShardedDataset._process_sample` (datasets/wds_dataset.py) for seeded
                        samples: pins oracle/wds.py and gdl_b200/wds_feeder.py to the reference itself (`--wds`).
* dynamic_mit_b0_golden.pt — logits of the REFERENCE's SegFormer with its DynamicMixTransformer encoder (`--dynamic`).
* unetpp_r18_golden.pt — seeded input / state_dict / logits / loss / selected gradients of
                        oracle/unetpp.py (resnet18, 3 bands, 5 classes, 64x64): a regression pin of
                        the oracle restatement (smp itself is not installable here, so its values
                        cannot be generated; see oracle/unetpp.py header).
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"


def _load_ref_tensors():
    spec = importlib.util.spec_from_file_location("ref_tensors", REF / "geo_deep_learning" / "utils" / "tensors.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def tensors_golden() -> None:
    ref = _load_ref_tensors()
    g = torch.Generator().manual_seed(20260925)
    cases = {}
    for c in (3, 4, 6):
        raw = torch.randint(0, 256, (2, c, 16, 16), generator=g, dtype=torch.uint8)
        mean = torch.rand(c, generator=g) * 0.5 + 0.2
        std = torch.rand(c, generator=g) * 0.3 + 0.1
        x = ref.normalization(raw.float())
        y = ref.standardization(x, mean.view(c, 1), std.view(c, 1))
        cases[f"c{c}"] = {"raw": raw, "mean": mean, "std": std, "normalized": x, "standardized": y}
    cases["norm_custom"] = {"in": torch.tensor([0.0, 255.0]),
                            "out": ref.normalization(torch.tensor([0.0, 255.0]), 0, 255, -1.0, 1.0)}
    torch.save(cases, OUT / "tensors_golden.pt")
    print("wrote tensors_golden.pt")


def build_seeded_r18():
    """Deterministic oracle weights (CPU RNG stream of this torch build) used by the golden file."""
    from oracle.unetpp import UnetPlusPlusOracle
    torch.manual_seed(7)
    m = UnetPlusPlusOracle("resnet18", 3, 5).train()
    # non-trivial BN affine parameters so their gradients are exercised
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
    return m


def unetpp_golden() -> None:
    m = build_seeded_r18()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.randint(0, 5, (2, 64, 64), generator=g)
    logits = m(x)
    loss = torch.nn.functional.cross_entropy(logits, t)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()
             if n in ("encoder.conv1.weight", "encoder.layer1.0.conv1.weight", "encoder.layer4.1.bn2.weight",
                      "decoder.blocks.x_0_0.conv1.0.weight", "decoder.blocks.x_0_3.conv2.1.bias",
                      "segmentation_head.0.weight", "segmentation_head.0.bias")}
    m.eval()
    with torch.no_grad():
        logits_eval = m(x)
    # weights are NOT stored (64 MB): they are regenerated from the seed by build_seeded_r18() below
    torch.save({"weight_checksum": sum(v.double().sum() for v in sd0.values() if v.is_floating_point()),
                "x": x, "target": t, "logits_train_sum": logits.detach().double().sum(),
                "logits_train_slice": logits.detach()[:, :, ::8, ::8].clone(), "loss": loss.detach(),
                "grad_norms": {k: v.norm() for k, v in grads.items()},
                "logits_eval_slice": logits_eval[:, :, ::8, ::8].clone()}, OUT / "unetpp_r18_golden.pt")
    print("wrote unetpp_r18_golden.pt")


def segformer_golden() -> None:
    """Outputs of the REFERENCE's SegFormerSegmentationModel (mit_b0, 3 bands, 5 classes, 64x64) loaded
    with oracle.segformer.init_state_dict(seed=0): pins oracle/segformer.py to the reference itself."""
    import torch.nn.functional as F
    from oracle import ref_shims
    from oracle import segformer as osf
    sd = osf.init_state_dict("mit_b0", 3, 5, seed=0)
    ref = ref_shims.reference_segformer("mit_b0", 3, 5)
    ref.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.randint(0, 5, (2, 64, 64), generator=g)
    ref.eval()
    with torch.no_grad():
        logits_eval = ref(x)
    ref.train()
    for m in ref.modules():  # stochastic layers off: parity is defined without dropout / drop-path
        if m.__class__.__name__ in ("DropPath", "Dropout2d", "Dropout"):
            m.eval()
    logits = ref(x)
    loss = F.cross_entropy(logits, t)
    loss.backward()
    names = ["encoder.patch_embed1.proj.weight", "encoder.block1.0.attn.q.weight", "encoder.block1.0.attn.sr.weight",
             "encoder.block2.1.mlp.dwconv.dwconv.weight", "encoder.block4.1.attn.kv.weight", "encoder.norm3.weight",
             "decoder.linear_c2.proj.weight", "decoder.linear_fuse.0.weight", "decoder.linear_fuse.1.bias",
             "decoder.linear_pred.weight"]
    params = dict(ref.named_parameters())
    torch.save({"x": x, "target": t, "logits_eval_slice": logits_eval[:, :, ::8, ::8].clone(),
                "logits_train_slice": logits.detach()[:, :, ::8, ::8].clone(), "loss": loss.detach(),
                "grad_norms": {n: params[n].grad.norm() for n in names},
                "grad_slices": {n: params[n].grad.flatten()[:64].clone() for n in names}},
               OUT / "segformer_b0_golden.pt")
    print("wrote segformer_b0_golden.pt")


def upernet_golden() -> None:
    """Outputs of the REFERENCE's MultiLevelNeck + UperNetDecoder + FCNHead + SegmentationHead (all torch-only,
    imported from /root/reference) loaded with oracle.upernet.init_state_dict(96, 64, 5, seed=1)."""
    import torch.nn as nn
    import torch.nn.functional as F
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from geo_deep_learning.models.decoders.upernet import UperNetDecoder
    from geo_deep_learning.models.heads.fcn_head import FCNHead
    from geo_deep_learning.models.heads.segmentation_head import SegmentationHead
    from geo_deep_learning.models.necks.multilevel_neck import MultiLevelNeck
    from oracle import upernet as ou
    e, ch, k = 96, 64, 5
    m = nn.Module()
    m.neck = MultiLevelNeck([e] * 4, [e] * 4, scales=[4, 2, 1, 0.5], norm_cfg={"type": "BN"}, act_cfg={"type": "ReLU"})
    m.decoder = UperNetDecoder([e] * 4, (1, 2, 3, 6), ch, align_corners=False, scale_modules=False)
    m.aux_head = FCNHead(e, ch, num_convs=1, num_classes=k, dropout_ratio=0.0)
    m.head = SegmentationHead(ch, k)
    m.load_state_dict(ou.init_state_dict(e, ch, k, seed=1))
    g = torch.Generator().manual_seed(2)
    feats = [torch.randn(2, e, 12, 12, generator=g) for _ in range(4)]
    t = torch.randint(0, k, (2, 168, 168), generator=g)
    m.train()
    f = m.neck(feats)
    out = F.interpolate(m.head(m.decoder(f)), size=(168, 168), mode="bilinear", align_corners=False)
    aux = F.interpolate(m.aux_head(f[-1]), size=(168, 168), mode="bilinear", align_corners=False)
    loss = F.cross_entropy(out, t) + 0.4 * F.cross_entropy(aux, t)
    loss.backward()
    names = ["neck.lateral_convs.0.conv.weight", "neck.convs.3.conv.bias", "decoder.psp_modules.2.1.conv.weight",
             "decoder.bottleneck.conv.weight", "decoder.fpn_convs.1.norm.weight", "decoder.fpn_bottleneck.conv.weight",
             "aux_head.cls_seg.weight", "head.conv.bias"]
    params = dict(m.named_parameters())
    torch.save({"feats": feats, "target": t, "out_slice": out.detach()[:, :, ::8, ::8].clone(),
                "aux_slice": aux.detach()[:, :, ::8, ::8].clone(), "loss": loss.detach(),
                "grad_slices": {n: params[n].grad.flatten()[:64].clone() for n in names}}, OUT / "upernet_golden.pt")
    print("wrote upernet_golden.pt")


def dofa_golden() -> None:
    """Feature maps of the REFERENCE's DOFAv2 encoder (imported from /root/reference on top of oracle/ref_shims.py's
    timm shim) loaded with oracle.dofa.init_state_dict(192, 4, 112, seed=2, ls_init=0.5); out_indices (1, 2, 3)."""
    from oracle import dofa as od, ref_shims
    e, depth, heads, img = 192, 4, 3, 112
    ref = ref_shims.reference_dofa(img, e, depth, heads, (1, 2, 3))
    sd = od.init_state_dict(e, depth, img, seed=2, ls_init=0.5)
    assert set(ref.state_dict()) == set(sd)
    ref.load_state_dict(sd)
    ref.eval()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 5, img, img, generator=g)
    wl = torch.tensor([0.49, 0.56, 0.665, 0.842, 1.61])
    with torch.no_grad():
        feats = ref(x, wl)
    torch.save({"x": x, "wavelengths": wl, "feats": [f.clone() for f in feats]}, OUT / "dofa_golden.pt")
    print("wrote dofa_golden.pt")


def dynamic_mit_golden() -> None:
    """Logits of the REFERENCE's SegFormerSegmentationModel(use_dynamic_encoder=True) (DynamicMixTransformer /
    DynamicChannelEmbed, mix_transformer.py:762-934; mit_b0, 5 classes, 64x64) for 3- and 6-band inputs, loaded with
    oracle.segformer.init_dynamic_state_dict(seed=4): pins oracle.segformer.dynamic_channel_embed to the reference."""
    from oracle import ref_shims
    from oracle import segformer as osf
    ref = ref_shims.reference_segformer("mit_b0", 3, 5, dynamic=True).eval()
    sd = osf.init_dynamic_state_dict("mit_b0", 5, seed=4)
    assert set(sd) == set(ref.state_dict())
    ref.load_state_dict(sd)
    g = torch.Generator().manual_seed(21)
    out = {}
    for c in (3, 6):
        x = torch.randn(2, c, 64, 64, generator=g)
        with torch.no_grad():
            y = ref(x)
        out[c] = {"x": x, "logits_slice": y[:, :, ::4, ::4].clone()}
    torch.save(out, OUT / "dynamic_mit_b0_golden.pt")
    print("wrote dynamic_mit_b0_golden.pt")


def wds_golden() -> None:
    """Outputs of the REFERENCE's own `ShardedDataset._process_sample` (datasets/wds_dataset.py:217-303) for seeded
    samples in the dofa / clay / unified formats.  `webdataset` and `pytorch_lightning` (imported at the top of that
    module, not installed here) are satisfied by empty stand-ins: `_process_sample` touches neither."""
    import json
    import tempfile
    import types

    import numpy as np
    for name in ("webdataset", "pytorch_lightning", "pytorch_lightning.utilities"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch_lightning.utilities"].rank_zero_only = lambda f: f
    sys.modules["webdataset"].WebDataset = object  # only named in a return annotation (:393)
    sys.modules["pytorch_lightning"].utilities = sys.modules["pytorch_lightning.utilities"]
    sys.path.insert(0, str(REF))
    from geo_deep_learning.datasets.wds_dataset import ShardedDataset

    sensor = "worldview-3-ortho_test"
    stats = {"statistics": {sensor: {"mean": [88.5, 97.25, 71.0, 120.75], "std": [41.0, 39.5, 37.25, 55.0],
                                     "band_count": 4, "patch_count": 3, "dtype": "uint8"}}}
    rng = np.random.default_rng(20261017)
    samples = []
    for i in range(3):
        meta = {"metadata": {"datetime": ["2021-07-04T15:30:00Z", "2019-12-30T03:05:00+00:00", "not a date"][i],
                             "coordinates_lat": [45.4215, -33.9, 0.0][i], "coordinates_lon": [-75.6972, 151.2, 180.0][i],
                             "red_wavelength": 0.66, "green_wavelength": 0.545, "blue_wavelength": 0.48,
                             "nir_wavelength": 0.8325}}
        samples.append({"__key__": f"patch_{i:04d}",
                        "image_patch.npy": rng.integers(0, 256, (4, 16, 16), dtype=np.uint8),
                        "label_patch.npy": rng.integers(0, 5, (1, 16, 16), dtype=np.uint8),
                        "metadata.json": meta})
    out = {"sensor": sensor, "stats": stats, "samples": samples, "outputs": {}}
    with tempfile.TemporaryDirectory() as td:
        sp = Path(td) / "stats.json"
        sp.write_text(json.dumps(stats))
        for mt in ("dofa", "clay", "unified"):
            ds = ShardedDataset(sensor, ["unused.tar"], 3, str(sp), model_type=mt, split="trn",
                                wavelength_keys=None if mt != "dofa" else ["red_wavelength", "green_wavelength",
                                                                           "blue_wavelength", "nir_wavelength"])
            out["outputs"][mt] = [ds._process_sample(dict(s)) for s in samples]
    torch.save(out, OUT / "wds_golden.pt")
    print("wrote wds_golden.pt")


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    tensors_golden()
    if "--all" in sys.argv:
        unetpp_golden()
        segformer_golden()
        upernet_golden()
        dofa_golden()
    if "--all" in sys.argv or "--wds" in sys.argv:
        wds_golden()
    if "--all" in sys.argv or "--dynamic" in sys.argv:
        dynamic_mit_golden()
