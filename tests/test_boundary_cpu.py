"""CPU: the C-ABI library builds/loads and exports what include/gdl_b200.h declares; host-side
logic (state_dict surface, error mapping, sharding) behaves like the reference's."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    h = (ROOT / "include" / "gdl_b200.h").read_text()
    return sorted(set(re.findall(r"\b(gdl_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from gdl_b200 import _build, _lib
    _build.build()
    lib = ctypes.CDLL(str(_lib.lib_path()))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in gdl_b200.h but not exported"


def test_binding_covers_every_declared_symbol():
    from gdl_b200 import _lib
    assert set(_declared()) == set(_lib.exported_symbols())
    assert _lib.load().gdl_version() == 100


def test_error_mapping_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    from gdl_b200 import _lib
    lib = _lib.load()
    d = _lib.ConvFwd()
    d.num_src = 0
    with pytest.raises(ValueError):
        _lib.check(lib.gdl_conv2d_nhwc_fwd(ctypes.byref(d), None))
    assert "num_src" in lib.gdl_last_error().decode()
    with pytest.raises(NotImplementedError):
        _lib.check(lib.gdl_seg_loss_fwd(1, 64, 1, 0, 10, 64, 0, 0, 1.0, 0.0, 0.0, 0, 0.0, 1e-7, 1, 1, None))


def test_cpu_tensors_are_rejected_loudly():
    from gdl_b200 import _lib
    with pytest.raises(_lib.GdlError):
        _lib.ptr(torch.zeros(4))


def test_state_dict_surface_matches_oracle():
    from gdl_b200.models.unetpp import UnetPlusPlus
    from oracle.unetpp import UnetPlusPlusOracle
    for enc, c, k in (("resnet18", 3, 5), ("resnet50", 4, 5), ("resnet34", 3, 2)):
        a = UnetPlusPlus(enc, in_channels=c, classes=k).state_dict()
        b = UnetPlusPlusOracle(enc, c, k).state_dict()
        assert list(a.keys()) == list(b.keys())
        for key in a:
            assert a[key].shape == b[key].shape, key


def test_unknown_encoder_raises_keyerror():
    from gdl_b200.models.unetpp import UnetPlusPlus
    with pytest.raises(KeyError):
        UnetPlusPlus("not_an_encoder")


def test_model_refuses_cpu_input():
    from gdl_b200.models.unetpp import UnetPlusPlus
    m = UnetPlusPlus("resnet18", in_channels=3, classes=2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 32, 32))


def test_checkpoints_round_trip_through_the_reference_loader(tmp_path):
    """`state_dict` drop-in, against the reference's OWN classes and loader (build container only): a Lightning-style
    checkpoint of the reference's SegFormerSegmentationModel loads strictly into gdl_b200's SegFormer through the
    reference's `load_weights_from_checkpoint` (utils/models.py:10-66, full and `load_parts` modes), and the product's
    state_dict loads strictly back into the reference model."""
    import importlib.util

    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("/root/reference not present (GPU box)")
    from gdl_b200.models.segformer import SegFormer
    ref = ref_shims.reference_segformer("mit_b1", 4, 3)
    spec = importlib.util.spec_from_file_location("ref_models_util", ref_shims.REF / "geo_deep_learning" / "utils" / "models.py")
    util = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(util)
    with torch.no_grad():
        for p in ref.parameters():
            p.normal_(0, 0.05)
    ckpt = tmp_path / "ref.ckpt"
    torch.save({"state_dict": {f"model.{k}": v for k, v in ref.state_dict().items()}, "epoch": 3}, ckpt)
    prod = SegFormer("mit_b1", in_channels=4, num_classes=3)
    assert util.load_weights_from_checkpoint(prod, str(ckpt)) is None  # strict load: every key and shape matches
    for k, v in ref.state_dict().items():
        assert torch.equal(prod.state_dict()[k], v), k
    prod2 = SegFormer("mit_b1", in_channels=4, num_classes=3)
    res = util.load_weights_from_checkpoint(prod2, str(ckpt), load_parts=["encoder"])
    assert not res.unexpected_keys and all(k.startswith("decoder.") for k in res.missing_keys)
    assert torch.equal(prod2.encoder.block2[1].attn.kv.weight, ref.state_dict()["encoder.block2.1.attn.kv.weight"])
    ref.load_state_dict(prod.state_dict())  # and back, strictly


def test_dofa_encoder_state_dict_matches_the_reference_class():
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("/root/reference not present (GPU box)")
    from gdl_b200.models.dofa import DOFAv2
    ref = ref_shims.reference_dofa(56, 64, 2, 4, out_indices=(0, 1))
    prod = DOFAv2("dofa_base", 56, 14, 64, 2, 4, out_indices=[0, 1])
    rs, ps = ref.state_dict(), prod.state_dict()
    assert set(rs) == set(ps)
    assert all(rs[k].shape == ps[k].shape for k in rs)
    prod.load_state_dict(rs)
    ref.load_state_dict(ps)


def test_reference_patch_first_conv_works_on_the_product_models(monkeypatch):
    """The reference adapts a 3-band (pretrained) stem to N bands with `patch_first_conv` (models/utils.py:140-181:
    weights cycled over the bands and scaled by 3/N).  The product models keep the stem as an ordinary nn.Conv2d, so the
    reference's own function applies unchanged, and the patched model computes what the oracle computes with the same
    patched tensors."""
    import importlib.util

    import cpu_kernel_emulation as emu
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("/root/reference not present (GPU box)")
    from gdl_b200.models.segformer import SegFormer
    from gdl_b200.models.unetpp import UnetPlusPlus
    from oracle import segformer as osf
    from oracle.unetpp import UnetPlusPlusOracle
    spec = importlib.util.spec_from_file_location("ref_model_utils", ref_shims.REF / "geo_deep_learning" / "models" / "utils.py")
    util = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(util)
    emu.install(monkeypatch)
    torch.manual_seed(0)
    m = UnetPlusPlus("resnet18", in_channels=3, classes=2, compute_dtype=torch.float32).eval()
    w3 = m.encoder.conv1.weight.detach().clone()
    util.patch_first_conv(m, 4, pretrained=True)
    assert m.encoder.conv1.weight.shape == (64, 4, 7, 7)
    assert torch.allclose(m.encoder.conv1.weight[:, 3], w3[:, 0] * 0.75) and torch.allclose(m.encoder.conv1.weight[:, 1], w3[:, 1] * 0.75)
    ora = UnetPlusPlusOracle("resnet18", 4, 2).eval()
    ora.load_state_dict(m.state_dict())
    x = torch.randn(1, 4, 32, 32)
    with torch.no_grad():
        assert torch.allclose(m(x), ora(x), atol=1e-4, rtol=1e-4)
    s = SegFormer("mit_b0", in_channels=3, num_classes=2, compute_dtype=torch.float32).eval()
    util.patch_first_conv(s, 6, pretrained=True)
    assert s.encoder.patch_embed1.proj.weight.shape[1] == 6
    x6 = torch.randn(1, 6, 32, 32)
    with torch.no_grad():
        assert torch.allclose(s(x6), osf.segformer_forward(s.state_dict(), x6, "mit_b0"), atol=1e-4, rtol=1e-4)


def test_every_kernel_waits_for_its_predecessor_before_touching_memory():
    """Programmatic dependent launch (csrc/common.cuh): with the "pdl" option every launch may become resident before the
    previous kernel of the stream has finished, so EVERY __global__ function must block in griddepcontrol.wait
    (GDL_PDL_ENTRY) before its first statement, and every launch must go through GDL_LAUNCH (the one place that attaches
    the attribute).  A kernel that skipped the wait would race with its producer; a raw <<< >>> launch is harmless but
    would silently opt out."""
    csrc = ROOT / "geo-deep-learning_b200" / "csrc"
    kernels = 0
    for f in sorted(csrc.glob("*.cu")) + sorted(csrc.glob("*.cuh")):
        text = f.read_text()
        code = re.sub(r"//[^\n]*", "", text)
        assert "<<<" not in code, f"{f.name}: raw <<< >>> launch (use GDL_LAUNCH)"
        for m in re.finditer(r"\b__global__\b", code):
            depth, i = 0, m.end()
            while True:  # the function body's opening brace: the first '{' outside parentheses
                c = code[i]
                if c == "(":
                    depth += 1
                elif c == ")":
                    depth -= 1
                elif c == "{" and depth == 0:
                    break
                elif c == ";" and depth == 0:
                    i = -1  # a declaration without body
                    break
                i += 1
            if i < 0:
                continue
            kernels += 1
            name = re.findall(r"(\w+)\s*\(", code[m.end():i])[-1]
            assert re.match(r"\s*GDL_PDL_ENTRY\(\);", code[i + 1:]), f"{f.name}: kernel {name} does not begin with GDL_PDL_ENTRY()"
    assert kernels >= 60
    common = (csrc / "common.cuh").read_text()
    assert "griddepcontrol.wait" in common and "cudaLaunchAttributeProgrammaticStreamSerialization" in common
