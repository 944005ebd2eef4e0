"""Thin, allocation-explicit Python wrappers over the C ABI (include/gdl_b200.h).

Activations are NHWC torch tensors of shape (N, H, W, C), 16-bit, whose last dim is
contiguous and whose pixel stride `stride(2)` may exceed C (channel slice of a wider buffer).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib as L


def _nhwc_src(t: torch.Tensor) -> tuple[int, int, int, int, int]:
    if t.dim() != 4:
        raise ValueError(f"NHWC activation expected 4 dims, got {tuple(t.shape)}")
    n, h, w, c = t.shape
    if t.stride(3) != 1:
        raise ValueError("NHWC activation: channel dim must be contiguous")
    ld = t.stride(2)
    if w > 1 and t.stride(1) != w * ld or (h > 1 and n > 1 and t.stride(0) != h * w * ld):
        raise ValueError(f"NHWC activation: non-uniform pixel stride {t.stride()} for shape {tuple(t.shape)}")
    return n, h, w, c, ld


def _fill_srcs(desc, srcs: Sequence[torch.Tensor]):
    if not 1 <= len(srcs) <= L.GDL_MAX_SRC:
        raise ValueError(f"between 1 and {L.GDL_MAX_SRC} sources expected, got {len(srcs)}")
    n0 = h0 = w0 = None
    for i, t in enumerate(srcs):
        n, h, w, c, ld = _nhwc_src(t)
        if i == 0:
            n0, h0, w0 = n, h, w
        elif (n, h, w) != (n0, h0, w0):
            raise ValueError("all concat sources must share N,H,W")
        desc.src[i].ptr = t.data_ptr()
        desc.src[i].channels = c
        desc.src[i].ld = ld
    desc.num_src = len(srcs)
    desc.N, desc.H, desc.W = n0, h0, w0
    return n0, h0, w0


def conv2d_fwd(srcs: Sequence[torch.Tensor], weight: torch.Tensor, cout: int, r: int, s: int,
               pad_h: int, pad_w: int, *, out: torch.Tensor | None = None,
               out_dtype: torch.dtype | None = None, bias: torch.Tensor | None = None,
               relu: bool = False) -> torch.Tensor:
    """gdl_conv2d_nhwc_fwd. `weight` is the packed [Cout][R][S][Ctot] 16-bit operand."""
    d = L.ConvFwd()
    n, h, w = _fill_srcs(d, srcs)
    dt = srcs[0].dtype
    ho, wo = h + 2 * pad_h - r + 1, w + 2 * pad_w - s + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=out_dtype or dt, device=srcs[0].device)
    d.Cout, d.R, d.S, d.pad_h, d.pad_w = cout, r, s, pad_h, pad_w
    d.weight = weight.data_ptr()
    d.dtype = L.dt_code(dt)
    d.out = out.data_ptr()
    d.out_dtype = L.dt_code(out.dtype)
    d.ldo = out.stride(2)
    d.bias = bias.data_ptr() if bias is not None else None
    d.relu = int(relu)
    L.check(L.load().gdl_conv2d_nhwc_fwd(C.byref(d), L.stream_ptr()))
    return out


def conv2d_wgrad(srcs: Sequence[torch.Tensor], dy: torch.Tensor, r: int, s: int, pad_h: int,
                 pad_w: int, dw: torch.Tensor) -> torch.Tensor:
    """gdl_conv2d_nhwc_wgrad: accumulates into fp32 dw [Cout][R][S][Ctot]."""
    d = L.ConvWgrad()
    _fill_srcs(d, srcs)
    d.Cout = dy.shape[3]
    d.R, d.S, d.pad_h, d.pad_w = r, s, pad_h, pad_w
    d.dy = dy.data_ptr()
    d.ld_dy = dy.stride(2)
    d.dtype = L.dt_code(srcs[0].dtype)
    if dw.dtype != torch.float32 or not dw.is_contiguous():
        raise ValueError("dw must be contiguous fp32")
    d.dw = dw.data_ptr()
    L.check(L.load().gdl_conv2d_nhwc_wgrad(C.byref(d), L.stream_ptr()))
    return dw


def pack_conv_weight(w_oihw: torch.Tensor, dtype: torch.dtype, transpose: bool = False,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 OIHW -> 16-bit [Cout][R][S][Cin] (or the dgrad operand [Cin][R][S][Cout], taps flipped)."""
    k, c, r, s = w_oihw.shape
    if w_oihw.dtype != torch.float32 or not w_oihw.is_contiguous():
        raise ValueError("weight must be contiguous fp32 OIHW")
    if out is None:
        shape = (c, r, s, k) if transpose else (k, r, s, c)
        out = torch.empty(shape, dtype=dtype, device=w_oihw.device)
    L.check(L.load().gdl_pack_conv_weight(L.ptr(w_oihw), L.ptr(out), k, c, r, s, int(transpose),
                                          L.dt_code(dtype), L.stream_ptr()))
    return out


def unpack_conv_wgrad(dw_krsc: torch.Tensor, out_oihw: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
    k, c, r, s = out_oihw.shape
    L.check(L.load().gdl_unpack_conv_wgrad(L.ptr(dw_krsc), L.ptr(out_oihw), k, c, r, s,
                                           int(accumulate), L.stream_ptr()))
    return out_oihw
