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
