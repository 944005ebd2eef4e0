"""CPU: pin the oracle against the reference's own known answers and committed golden vectors."""
from pathlib import Path

import pytest
import torch

from oracle import losses as olosses
from oracle import tensors as otensors
from oracle.unetpp import UnetPlusPlusOracle, decoder_plan

GOLD = Path(__file__).parent / "golden"


# --- the reference's own known-answer tests (tests/test_utils_tensors.py:14-50), restated -------------
def test_normalization_simple_range():
    t = torch.tensor([[0.0, 127.5, 255.0]])
    assert torch.allclose(otensors.normalization(t, 0, 255, 0.0, 1.0), torch.tensor([[0.0, 0.5, 1.0]]), atol=1e-6)


def test_normalization_custom_range():
    t = torch.tensor([0.0, 255.0])
    assert torch.allclose(otensors.normalization(t, 0, 255, -1.0, 1.0), torch.tensor([-1.0, 1.0]), atol=1e-6)


def test_standardization_basic():
    t = torch.tensor([[[[1.0, 2.0], [3.0, 4.0]]]])
    mean, std = torch.tensor([2.5]), torch.tensor([1.118034])
    exp = (t - mean.view(1, 1, 1, 1)) / std.view(1, 1, 1, 1)
    assert torch.allclose(otensors.standardization(t, mean, std), exp, atol=1e-6)


# --- golden vectors produced by the reference module itself (oracle/make_golden.py) -------------------
@pytest.mark.parametrize("c", [3, 4, 6])
def test_tensors_match_reference_golden(c):
    g = torch.load(GOLD / "tensors_golden.pt")[f"c{c}"]
    x = otensors.normalization(g["raw"].float())
    assert torch.equal(x, g["normalized"])
    y = otensors.standardization(x, g["mean"].view(c, 1), g["std"].view(c, 1))
    assert torch.equal(y, g["standardized"])
    # per-sample path used by the WebDataset pipeline
    y0 = otensors.patch_normalise(g["raw"][0], g["mean"], g["std"])
    assert torch.allclose(y0, g["standardized"][0], atol=1e-6)


# --- UNet++ restatement: shapes pinned by the notebook's parameter count ------------------------------
def test_unetpp_param_count_matches_notebook():
    # notebooks/00_quickstart.ipynb:572-576 records "26.1 M" for UnetPlusPlus-resnet34 / 3 bands / 2 classes
    m = UnetPlusPlusOracle("resnet34", 3, 2)
    assert sum(p.numel() for p in m.parameters()) == 26_078_754


def test_unetpp_param_counts_baseline_configs():
    assert sum(p.numel() for p in UnetPlusPlusOracle("resnet50", 4, 5).parameters()) == 48_989_461
    assert sum(p.numel() for p in UnetPlusPlusOracle("resnet18", 3, 5).parameters()) == 15_971_029


def test_unetpp_decoder_plan_r50():
    plan = decoder_plan((64, 256, 512, 1024, 2048))
    assert plan["x_0_0"] == (2048, 1024, 256)
    assert plan["x_1_1"] == (1024, 512, 512)
    assert plan["x_0_3"] == (64, 256, 32)
    assert plan["x_0_4"] == (32, 0, 16)
    assert len(plan) == 11


def test_unetpp_rejects_non_multiple_of_32():
    m = UnetPlusPlusOracle("resnet18", 3, 2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 48, 40))


def test_unetpp_golden_regression():
    from oracle.make_golden import build_seeded_r18
    g = torch.load(GOLD / "unetpp_r18_golden.pt")
    m = build_seeded_r18()
    chk = sum(v.double().sum() for v in m.state_dict().values() if v.is_floating_point())
    if abs(float(chk) - float(g["weight_checksum"])) > 1e-6 * abs(float(g["weight_checksum"])):
        pytest.skip("CPU RNG stream differs from the build container: golden weights cannot be regenerated")
    logits = m(g["x"])
    loss = torch.nn.functional.cross_entropy(logits, g["target"])
    assert torch.allclose(logits[:, :, ::8, ::8], g["logits_train_slice"], atol=1e-4, rtol=1e-4)
    assert torch.allclose(loss, g["loss"], atol=1e-5)


# --- losses ----------------------------------------------------------------------------------------
def test_dice_multiclass_perfect_prediction_is_zero():
    t = torch.randint(0, 3, (2, 8, 8))
    logits = torch.nn.functional.one_hot(t, 3).permute(0, 3, 1, 2).float() * 50.0
    assert olosses.dice_loss(logits, t, "multiclass").item() < 1e-5


def test_dice_absent_class_contributes_zero():
    t = torch.zeros(1, 4, 4, dtype=torch.long)  # only class 0 present, K = 3
    logits = torch.zeros(1, 3, 4, 4)
    # class 0: p = 1/3 everywhere: I = 16/3, C = 16/3 + 16 -> dice = 0.5, loss 0.5; others masked -> mean = 0.5/3
    assert abs(olosses.dice_loss(logits, t, "multiclass").item() - 0.5 / 3) < 1e-6


def test_soft_ce_reduces_to_ce_without_smoothing():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 8, 8, generator=g)
    t = torch.randint(0, 5, (2, 8, 8), generator=g)
    assert torch.allclose(olosses.soft_ce_loss(x, t, 0.0), torch.nn.functional.cross_entropy(x, t), atol=1e-6)
    # torch's label_smoothing formula is the same one
    assert torch.allclose(olosses.soft_ce_loss(x, t, 0.1, None),
                          torch.nn.functional.cross_entropy(x, t, label_smoothing=0.1), atol=1e-6)


# --- SegFormer restatement: pinned to the reference's own modules ------------------------------------
def _close(a, b, rel=5e-6):
    """max |a-b| <= rel * max |b|  (fp32 re-association noise between two summation orders)"""
    return (a - b).abs().max().item() <= rel * b.abs().max().item() + 1e-12


def _segformer_train_outputs(sd, g):
    from oracle import segformer as osf
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd.items()}
    logits = osf.segformer_forward(sd, g["x"], "mit_b0", training=True)
    loss = torch.nn.functional.cross_entropy(logits, g["target"])
    loss.backward()
    return sd, logits, loss


def test_segformer_oracle_matches_reference_golden():
    """tests/golden/segformer_b0_golden.pt was produced by the REFERENCE's SegFormerSegmentationModel."""
    from oracle import segformer as osf
    g = torch.load(GOLD / "segformer_b0_golden.pt")
    sd0 = osf.init_state_dict("mit_b0", 3, 5, seed=0)
    with torch.no_grad():
        ev = osf.segformer_forward(sd0, g["x"], "mit_b0", training=False)
    assert _close(ev[:, :, ::8, ::8], g["logits_eval_slice"])
    sd, logits, loss = _segformer_train_outputs(sd0, g)
    assert _close(logits[:, :, ::8, ::8], g["logits_train_slice"])
    assert torch.allclose(loss, g["loss"], atol=1e-6)
    for n, want in g["grad_slices"].items():
        assert _close(sd[n].grad.flatten()[:64], want, 1e-4), n


def test_segformer_oracle_matches_reference_import():
    """When /root/reference is mounted (build container), compare with the live reference modules."""
    from oracle import ref_shims
    from oracle import segformer as osf
    if not ref_shims.available():
        pytest.skip("/root/reference not present (GPU box): covered by the golden file")
    sd0 = osf.init_state_dict("mit_b2", 4, 5, seed=3)
    ref = ref_shims.reference_segformer("mit_b2", 4, 5)
    assert set(ref.state_dict().keys()) == set(sd0.keys())
    ref.load_state_dict(sd0)
    assert sum(p.numel() for p in ref.parameters()) == 27_350_469 + 3136  # +3 136 per extra band
    ref.eval()
    x = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        assert _close(osf.segformer_forward(sd0, x, "mit_b2"), ref(x))


def test_dynamic_mix_transformer_oracle_matches_reference_golden():
    """tests/golden/dynamic_mit_b0_golden.pt: outputs of the REFERENCE's SegFormer with its DynamicMixTransformer encoder."""
    from oracle import segformer as osf
    g = torch.load(GOLD / "dynamic_mit_b0_golden.pt")
    sd = osf.init_dynamic_state_dict("mit_b0", 5, seed=4)
    for c, case in g.items():
        with torch.no_grad():
            y = osf.segformer_forward(sd, case["x"], "mit_b0")
        assert _close(y[:, :, ::4, ::4], case["logits_slice"]), c


def test_dynamic_mix_transformer_oracle_matches_reference_import():
    from oracle import ref_shims
    from oracle import segformer as osf
    if not ref_shims.available():
        pytest.skip("/root/reference not present (GPU box): covered by the golden file")
    ref = ref_shims.reference_segformer("mit_b1", 3, 4, dynamic=True).eval()
    sd = osf.init_dynamic_state_dict("mit_b1", 4, seed=9)
    assert set(ref.state_dict().keys()) == set(sd.keys())
    ref.load_state_dict(sd)
    for c in (1, 4, 8):
        x = torch.randn(1, c, 64, 64, generator=torch.Generator().manual_seed(c))
        with torch.no_grad():
            assert _close(osf.segformer_forward(sd, x, "mit_b1"), ref(x)), c


# --- MultiLevelNeck + UperNet + heads restatement: pinned to the reference's own modules -----------------
def test_upernet_oracle_matches_reference_golden():
    from oracle import upernet as ou
    g = torch.load(GOLD / "upernet_golden.pt")
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in ou.init_state_dict(96, 64, 5, seed=1).items()}
    out, aux = ou.upernet_forward(sd, g["feats"], (168, 168), training=True)
    loss = torch.nn.functional.cross_entropy(out, g["target"]) + 0.4 * torch.nn.functional.cross_entropy(aux, g["target"])
    loss.backward()
    assert _close(out[:, :, ::8, ::8], g["out_slice"]) and _close(aux[:, :, ::8, ::8], g["aux_slice"])
    assert torch.allclose(loss, g["loss"], atol=1e-6)
    for n, want in g["grad_slices"].items():
        assert _close(sd[n].grad.flatten()[:64], want, 1e-4), n


# --- DOFA-v2 encoder restatement: pinned to the reference's DOFAv2 (timm Block restated, see oracle/dofa.py) ----
def test_dofa_oracle_matches_reference_golden():
    from oracle import dofa as od
    g = torch.load(GOLD / "dofa_golden.pt")
    sd = od.init_state_dict(192, 4, 112, seed=2, ls_init=0.5)
    feats = od.dofa_forward(sd, g["x"], g["wavelengths"], 192, 4, 3, out_indices=(1, 2, 3))
    assert len(feats) == 3
    for a, b in zip(feats, g["feats"]):
        assert a.shape == b.shape == (2, 192, 8, 8) and _close(a, b, 2e-5)


def test_dofa_oracle_matches_reference_import():
    from oracle import dofa as od, ref_shims
    if not ref_shims.available():
        pytest.skip("/root/reference not present (GPU box): covered by the golden file")
    sd = od.init_state_dict(96, 2, 56, seed=7, ls_init=1.0)
    ref = ref_shims.reference_dofa(56, 96, 2, 3, (0, 1))
    assert set(ref.state_dict()) == set(sd)
    ref.load_state_dict(sd)
    ref.eval()
    x = torch.randn(1, 3, 56, 56, generator=torch.Generator().manual_seed(0))
    wl = torch.tensor([0.665, 0.56, 0.49])
    with torch.no_grad():
        want = ref(x, wl)
    for a, b in zip(od.dofa_forward(sd, x, wl, 96, 2, 3, out_indices=(0, 1)), want):
        assert _close(a, b, 2e-5)


def test_dofa_convert_patch_to_16_matches_reference_import():
    """`convert_patch_to_16=True` (dofa_v2.py:168-176,220): generated 14x14 kernels resampled to 16x16, stride 16."""
    from oracle import dofa as od, ref_shims
    if not ref_shims.available():
        pytest.skip("/root/reference not present (GPU box)")
    sd = od.init_state_dict(96, 2, 64, seed=7, ls_init=1.0)
    sd["pos_embed"] = od.sincos_2d(96, 4).unsqueeze(0)  # (64 // 16)^2 patches
    ref = ref_shims.reference_dofa(64, 96, 2, 3, (0, 1), convert_patch_to_16=True)
    assert ref.pos_embed.shape == sd["pos_embed"].shape
    ref.load_state_dict(sd)
    ref.eval()
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    wl = torch.tensor([0.665, 0.56, 0.49])
    with torch.no_grad():
        want = ref(x, wl)
    for a, b in zip(od.dofa_forward(sd, x, wl, 96, 2, 3, out_indices=(0, 1), convert_to_16=True), want):
        assert a.shape == b.shape == (1, 96, 4, 4) and _close(a, b, 2e-5)
