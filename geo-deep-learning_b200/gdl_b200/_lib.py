"""ctypes binding of libgdlb200.so (C ABI declared in include/gdl_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, the
product path raises.  (oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

GDL_OK, GDL_ERR_INVALID, GDL_ERR_UNSUPPORTED, GDL_ERR_CUDA = 0, 1, 2, 3
GDL_BF16, GDL_F32, GDL_F16 = 0, 1, 2
GDL_MAX_SRC = 6

_LIB_PATH = Path(__file__).resolve().parent / "libgdlb200.so"
_lib = None


class GdlError(RuntimeError):
    pass


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("channels", C.c_int), ("ld", C.c_int)]


class ConvFwd(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("num_src", C.c_int),
        ("src", Src * GDL_MAX_SRC),
        ("Cout", C.c_int),
        ("R", C.c_int), ("S", C.c_int), ("pad_h", C.c_int), ("pad_w", C.c_int),
        ("weight", C.c_void_p),
        ("dtype", C.c_int),
        ("out", C.c_void_p),
        ("out_dtype", C.c_int),
        ("ldo", C.c_int),
        ("bias", C.c_void_p),
        ("relu", C.c_int),
        ("residual", C.c_void_p),
        ("res_dtype", C.c_int),
        ("ldr", C.c_int),
        ("oscale", C.c_void_p),
        ("w_ld", C.c_int),
        ("w_rows", C.c_int),
        ("w_rows_per_img", C.c_int),
        ("w_mn_major", C.c_int),
        ("groups", C.c_int),
        ("g_src_stride", C.c_int),
        ("g_w_stride", C.c_int),
        ("g_out_stride", C.c_int),
        ("bn_sums", C.c_void_p),
        ("bn_pivot", C.c_void_p),
    ]


class ConvWgrad(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("num_src", C.c_int),
        ("src", Src * GDL_MAX_SRC),
        ("Cout", C.c_int),
        ("R", C.c_int), ("S", C.c_int), ("pad_h", C.c_int), ("pad_w", C.c_int),
        ("dy", C.c_void_p),
        ("ld_dy", C.c_int),
        ("dtype", C.c_int),
        ("dw", C.c_void_p),
        ("dw_ld", C.c_int),
        ("dw_img_stride", C.c_longlong),
        ("batched", C.c_int),
        ("groups", C.c_int),
        ("g_src_stride", C.c_int),
        ("g_dy_stride", C.c_int),
        ("g_dw_stride", C.c_int),
    ]


class Repack(C.Structure):
    """gdl_repack_t: one packed operand of the batched refresh (gdl_repack_weights)"""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("Cout", C.c_int), ("Cin", C.c_int), ("R", C.c_int),
                ("S", C.c_int), ("mode", C.c_int), ("dst_ld", C.c_int)]


class Lowres(C.Structure):
    """gdl_lowres_t: one low-resolution source of gdl_bilinear_sum_fwd"""
    _fields_ = [("ptr", C.c_void_p), ("H", C.c_int), ("W", C.c_int), ("ld", C.c_longlong)]


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise GdlError(
            f"{_LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(str(_LIB_PATH))
    lib.gdl_last_error.restype = C.c_char_p
    lib.gdl_version.restype = C.c_int
    lib.gdl_query_workspace_bytes.restype = C.c_longlong
    lib.gdl_p2p_exchange_bytes.restype = C.c_longlong
    lib.gdl_p2p_exchange_bytes.argtypes = [C.c_int, C.c_int]
    _declare(lib)
    lib.gdl_set_option(b"pdl", PDL)
    _lib = lib
    return lib


# Programmatic dependent launch of every kernel (include/gdl_b200.h, option "pdl"; csrc/common.cuh): the host owns the
# switch so that what ran can be reported (bench.py config.pdl).  GDL_PDL=0/1 overrides the default.
PDL = int(os.environ.get("GDL_PDL", "0") != "0")


def set_pdl(on: bool) -> None:
    global PDL
    PDL = int(bool(on))
    load().gdl_set_option(b"pdl", PDL)


_VP, _I, _LL, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# argument types of every entry point in include/gdl_b200.h (explicit: long long / float args
# would otherwise be passed as 32-bit ints by ctypes)
_SIGS = {
    "gdl_conv2d_nhwc_fwd": [_VP, _VP],
    "gdl_conv2d_nhwc_wgrad": [_VP, _VP],
    "gdl_conv2d_bn_fusable": [_VP, _VP],
    "gdl_pack_conv_weight": [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_unpack_conv_wgrad": [_VP, _VP, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_widen_conv_weight": [_VP, _VP, _I, _I, _I, _I, _I, _VP],
    "gdl_fold_widened_wgrad": [_VP, _I, _I, _VP, _I, _I, _I, _I, _I, _VP],
    "gdl_normalize_to_nhwc": [_VP, _I, _VP, _I, _LL, _LL, _LL, _I, _I, _VP, _VP, _F, _VP],
    "gdl_augment_normalize": [_VP, _I, _VP, _I, _VP, _VP, _I, _VP, _LL, _LL, _LL, _I, _I, _VP, _VP, _F, _VP],
    "gdl_dropout2d_apply": [_VP, _LL, _VP, _VP, _LL, _I, _LL, _LL, _I, _VP],
    "gdl_im2col_nhwc": [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_col2im_nhwc": [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_bn_stats": [_VP, _I, _LL, _I, _I, _VP, _VP, _VP],
    "gdl_bn_finalize": [_VP, _VP, _LL, _I, _VP, _VP, _F, _F, _VP, _VP, _VP, _VP, _VP, _VP, _VP],
    "gdl_bn_eval_coeffs": [_I, _VP, _VP, _VP, _VP, _F, _VP, _VP, _VP],
    "gdl_bn_apply": [_VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _I, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_grad_gather": [_I, _VP, _VP, _VP, _VP, _I, _VP, _I, _VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _VP],
    "gdl_bn_bwd_apply": [_VP, _I, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _I, _LL, _LL, _I, _VP],
    "gdl_bn_param_grads": [_VP, _I, _VP, _VP, _I, _VP],
    "gdl_maxpool3x3s2_fwd": [_VP, _I, _VP, _I, _VP, _I, _I, _I, _I, _I, _VP],
    "gdl_maxpool3x3s2_bwd": [_VP, _I, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_seg_loss_fwd": [_VP, _I, _VP, _I, _LL, _I, _LL, _I, _F, _F, _F, _I, _F, _F, _VP, _VP, _VP],
    "gdl_seg_loss_bwd": [_VP, _I, _VP, _I, _LL, _I, _LL, _I, _F, _F, _F, _I, _F, _F, _VP, _VP, _VP, _I, _I, _VP],
    "gdl_argmax_classes": [_VP, _I, _LL, _I, _F, _VP, _VP],
    "gdl_upsample_ce_fwd": [_VP, _I, _I, _I, _I, _I, _I, _VP, _I, _I, _LL, _I, _F, _F, _F, _I, _F, _F, _VP, _VP, _VP],
    "gdl_upsample_ce_bwd": [_VP, _I, _I, _I, _I, _I, _I, _VP, _I, _I, _LL, _I, _F, _F, _F, _I, _F, _F, _VP, _VP, _VP, _I, _I, _VP],
    "gdl_upsample_argmax": [_VP, _I, _I, _I, _I, _I, _I, _I, _F, _VP, _VP],
    "gdl_argmax_confusion": [_VP, _I, _LL, _LL, _I, _F, _VP, _I, _LL, _I, _VP, _VP, _VP],
    "gdl_adam_step": [_VP, _VP, _VP, _VP, _LL, _F, _F, _F, _F, _F, _I, _VP, _VP],
    "gdl_adam_step_dev": [_VP, _VP, _VP, _VP, _LL, _F, _F, _F, _F, _F, _VP, _VP, _VP, _VP],
    "gdl_grad_clip_coef": [_VP, _LL, _F, _VP, _VP, _VP],
    "gdl_device_info": [_VP, _VP, _VP, _VP],
    "gdl_set_option": [C.c_char_p, _LL],
    "gdl_layernorm_fwd": [_VP, _I, _LL, _VP, _VP, _F, _VP, _I, _LL, _VP, _VP, _LL, _I, _VP],
    "gdl_layernorm_bwd": [_VP, _I, _LL, _VP, _I, _LL, _VP, _VP, _VP, _VP, _LL, _VP, _LL, _VP, _I, _LL, _VP, _LL, _I, _VP],
    "gdl_softmax_fwd": [_VP, _LL, _F, _VP, _LL, _I, _LL, _I, _I, _VP],
    "gdl_softmax_bwd": [_VP, _LL, _VP, _LL, _F, _VP, _LL, _I, _LL, _I, _I, _VP],
    "gdl_sra_attention_fwd": [_VP, _LL, _VP, _LL, _VP, _LL, _VP, _LL, _I, _I, _I, _I, _I, _F, _I, _VP],
    "gdl_sra_attention_bwd": [_VP, _LL, _VP, _LL, _VP, _LL, _VP, _LL, _VP, _LL, _I, _I, _I, _I, _I, _F, _I, _VP],
    "gdl_mha_flash_fwd": [_VP, _LL, _VP, _LL, _VP, _LL, _VP, _LL, _I, _I, _I, _F, _I, _VP],
    "gdl_dwconv3x3_gelu_fwd": [_VP, _I, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP],
    "gdl_dwconv3x3_gelu_bwd": [_VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _VP],
    "gdl_bilinear_fwd": [_VP, _LL, _VP, _LL, _I, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_bilinear_bwd": [_VP, _LL, _VP, _LL, _I, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_bilinear_sum_fwd": [_VP, _LL, _I, _VP, _VP, _LL, _I, _I, _I, _I, _I, _VP],
    "gdl_cast_f32": [_VP, _VP, _I, _LL, _VP],
    "gdl_vit_assemble_tokens": [_VP, _I, _VP, _VP, _VP, _I, _I, _I, _VP],
    "gdl_vit_extract_feature": [_VP, _VP, _I, _I, _I, _I, _VP],
    "gdl_vit_feature_grad": [_VP, _I, _VP, _I, _I, _I, _I, _VP],
    "gdl_gelu_fwd": [_VP, _VP, _I, _LL, _VP],
    "gdl_gelu_bwd": [_VP, _VP, _VP, _I, _LL, _VP],
    "gdl_layerscale_add": [_VP, _VP, _I, _VP, _VP, _LL, _VP, _LL, _I, _VP],
    "gdl_layerscale_bwd": [_VP, _VP, _I, _VP, _VP, _LL, _VP, _VP, _LL, _I, _VP],
    "gdl_channel_pool_fwd": [_VP, _VP, _VP, _VP, _I, _LL, _I, _I, _VP],
    "gdl_channel_pool_bwd": [_VP, _VP, _VP, _VP, _VP, _I, _LL, _I, _I, _VP],
    "gdl_relu_bwd": [_VP, _VP, _VP, _I, _LL, _VP],
    "gdl_adaptive_avgpool_fwd": [_VP, _LL, _VP, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_adaptive_avgpool_bwd": [_VP, _VP, _I, _I, _I, _I, _I, _I, _VP],
    "gdl_add_nhwc": [_VP, _LL, _VP, _LL, _VP, _LL, _I, _LL, _I, _VP],
    "gdl_set_workspace": [_VP, _LL, _VP],
    "gdl_repack_weights": [_VP, _VP, _I, _I, _I, _VP],
    "gdl_p2p_allreduce_sums": [_VP, _I, _VP, _I, _I, _I, _VP, _VP],
}


def exported_symbols() -> list[str]:
    return ["gdl_last_error", "gdl_version", "gdl_query_workspace_bytes", "gdl_p2p_exchange_bytes", *_SIGS.keys()]


def _declare(lib) -> None:
    for name, args in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol: fail loudly
        fn.argtypes = args
        fn.restype = C.c_int


def check(status: int) -> None:
    if status == GDL_OK:
        return
    msg = load().gdl_last_error().decode("utf-8", "replace")
    if status == GDL_ERR_INVALID:
        raise ValueError(msg)
    if status == GDL_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise GdlError(msg)


# ---- deterministic-reduction workspace (include/gdl_b200.h: gdl_set_workspace) ------------------------------------
# One caller-owned scratch buffer per device, registered before the first launch on that device: with it every
# cross-block sum of the library is formed in a fixed order (bit-reproducible steps; a CUDA-graph replay equals the
# eager launch).  GDL_DETERMINISTIC=0 leaves it unregistered (fp32 atomics, arrival order).
_WORKSPACES: dict[int, torch.Tensor] = {}


def _register_workspace(dev: int) -> None:
    import os
    if os.environ.get("GDL_DETERMINISTIC", "1") == "0":
        _WORKSPACES[dev] = None
        return
    if torch.cuda.is_current_stream_capturing():
        raise GdlError("the reduction workspace must be registered before CUDA-graph capture: run one eager call first")
    lib = load()
    nbytes = int(lib.gdl_query_workspace_bytes())
    buf = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", dev))
    check(lib.gdl_set_workspace(C.c_void_p(buf.data_ptr()), nbytes, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    _WORKSPACES[dev] = buf


def workspace_tensor():
    """the reduction workspace registered for the current device (uint8 tensor), or None"""
    return _WORKSPACES.get(torch.cuda.current_device())


def deterministic() -> bool:
    """True when the current device has a registered reduction workspace (ordered sums instead of fp32 atomics)."""
    return workspace_tensor() is not None


def stream_ptr() -> C.c_void_p:
    """current torch stream as a cudaStream_t; every kernel wrapper passes through here, so this is also where the
    per-device workspace is registered (once)"""
    dev = torch.cuda.current_device()
    if dev not in _WORKSPACES:
        _register_workspace(dev)
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_DT = {torch.bfloat16: GDL_BF16, torch.float32: GDL_F32, torch.float16: GDL_F16}


def dt_code(dtype: torch.dtype) -> int:
    try:
        return _DT[dtype]
    except KeyError:
        raise ValueError(f"unsupported dtype {dtype}") from None


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise GdlError("libgdlb200 operates on CUDA tensors only (no CPU fallback)")
    return C.c_void_p(t.data_ptr())
