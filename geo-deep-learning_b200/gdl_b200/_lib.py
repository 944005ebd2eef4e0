"""ctypes binding of libgdlb200.so (C ABI declared in include/gdl_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, the
product path raises.  (oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
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
    ]


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
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == GDL_OK:
        return
    msg = load().gdl_last_error().decode("utf-8", "replace")
    if status == GDL_ERR_INVALID:
        raise ValueError(msg)
    if status == GDL_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise GdlError(msg)


def stream_ptr() -> C.c_void_p:
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
