"""Thin, allocation-explicit Python wrappers over the C ABI (include/gdl_b200.h).

Activations are NHWC torch tensors of shape (N, H, W, C), 16-bit, whose last dim is
contiguous and whose pixel stride `stride(2)` may exceed C (channel slice of a wider buffer).
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Sequence

import torch

from . import _lib as L


def _ck(status: int) -> None:
    """check the status of a C-ABI call that launched (at least) one kernel"""
    L.check(status)
    _count()


# ---------------------------------------------------------------------------------------------
# instrumentation: launch counter (bench.py's gpu_launches) and per-launch conv timing (roofline)
# ---------------------------------------------------------------------------------------------
_LAUNCHES = 0
_PROFILER = None


def _count(n: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def launch_count() -> int:
    return _LAUNCHES


def reset_launch_count() -> None:
    global _LAUNCHES
    _LAUNCHES = 0


class ConvProfiler:
    """CUDA-event pair around every tensor-core conv launch on the launching stream."""

    def __init__(self) -> None:
        self.records: list[tuple[str, float, torch.cuda.Event, torch.cuda.Event]] = []
        self.descs: list[str] = []

    def begin(self) -> torch.cuda.Event:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, kernel: str, flops: float, e0: torch.cuda.Event, desc: str = "") -> None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append((kernel, flops, e0, e1))
        self.descs.append(desc)

    def dump_table(self, path: str, steps: int) -> None:
        """per-launch table of the LAST profiled step: kernel, shape, ms, TFLOP/s (sorted by time)."""
        import json
        torch.cuda.synchronize()
        n = len(self.records) // steps
        rows = []
        for (kernel, flops, e0, e1), desc in zip(self.records[-n:], self.descs[-n:]):
            ms = e0.elapsed_time(e1)
            rows.append({"kernel": kernel, "shape": desc, "ms": round(ms, 4), "gflop": round(flops / 1e9, 2),
                         "tflops": round(flops / ms / 1e9, 1) if ms > 0 else 0})
        rows.sort(key=lambda r: -r["ms"])
        with open(path, "w") as f:
            json.dump(rows, f, indent=0)

    def summary(self, steps: int = 1) -> dict:
        torch.cuda.synchronize()
        out: dict[str, dict] = {}
        for kernel, flops, e0, e1 in self.records:
            d = out.setdefault(kernel, {"ms": 0.0, "flops": 0.0, "launches": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["launches"] += 1
        for d in out.values():
            d["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
            d["ms_per_step"] = d["ms"] / steps
            d["launches_per_step"] = d["launches"] / steps
        for k in ("conv_fwd_kernel", "conv_wgrad_kernel"):
            out.setdefault(k, {"ms": 0.0, "flops": 0.0, "launches": 0, "tflops": 0.0, "ms_per_step": 0.0,
                               "launches_per_step": 0.0})
        return out


def set_conv_profiler(p: ConvProfiler | None) -> None:
    global _PROFILER
    _PROFILER = p


# host-side switches (the library's own are in gdl_set_option); pixel_pack: run 16/32-channel 3x3 convs pixel-packed
# sra_fused: SegFormer attention forward as ONE kernel (gdl_sra_attention_fwd) where its shape limits allow
# mha_flash: the DOFA encoder's (forward-only) self-attention as ONE kernel (gdl_mha_flash_fwd, keys streamed)
_HOST_OPTS = {"pixel_pack": int(os.environ.get("GDL_PIXEL_PACK", "1")), "sra_fused": int(os.environ.get("GDL_SRA_FUSED", "1")),
              "mha_flash": int(os.environ.get("GDL_MHA_FLASH", "1")),
              # attn_wgrad_grouped: dV / dK of all attention heads as one grouped wgrad launch each (instead of 2 x heads launches)
              "attn_wgrad_grouped": int(os.environ.get("GDL_ATTN_WGRAD_GROUPED", "1")),
              # fused_head: the fused trainer never materialises the upsampled (N,H,W,K) logits of SegFormer / UperNet heads:
              # gdl_upsample_ce_fwd / _bwd interpolate them on the fly (training); eval masks come from gdl_upsample_argmax
              "fused_head": int(os.environ.get("GDL_FUSED_HEAD", "1")),
              # bn_fused: a conv followed by a training-mode BatchNorm produces that layer's batch sums itself (from the
              # epilogue's staged tile) instead of a separate pass over its output (gdl_conv_fwd_t.bn_sums)
              "bn_fused": int(os.environ.get("GDL_BN_FUSED", "1")),
              # p2p_syncbn: SyncBN statistics exchanged over NVLink peer memory (gdl_p2p_allreduce_sums) instead of NCCL
              "p2p_syncbn": int(os.environ.get("GDL_P2P_SYNCBN", "1")),
              # overlap_allreduce: the flat gradient is all-reduced in 3 buckets, each started as soon as its layers' backward
              # is done (UNet++ route of the fused trainer); 0 = one all-reduce after the backward
              "overlap_allreduce": int(os.environ.get("GDL_OVERLAP_ALLREDUCE", "1")),
              # decoder_folded: SegFormer's linear_fuse 1x1 conv is applied in front of the (linear) bilinear resizes — per
              # level at its own resolution, composed with linear_c{l} — instead of on the concatenated resized maps
              # (models/segformer.py: _decoder_folded_fwd); 0 = the reference's op order
              "decoder_folded": int(os.environ.get("GDL_DECODER_FOLDED", "1"))}


def option(name: str) -> int:
    if name == "pdl":
        return L.PDL
    return _HOST_OPTS[name]


def set_option(name: str, value: int) -> None:
    """runtime tuning switch of the library (see gdl_set_option in include/gdl_b200.h) or of the host engine"""
    if name in _HOST_OPTS:
        _HOST_OPTS[name] = int(value)
        return
    if name == "pdl":
        L.set_pdl(bool(value))
        return
    L.check(L.load().gdl_set_option(name.encode(), int(value)))


class P2PExchange:
    """SyncBatchNorm statistics over NVLink peer memory (gdl_p2p_allreduce_sums): one symmetric exchange buffer per rank,
    mapped by every peer through torch.distributed._symmetric_memory.  Construct on every rank of `group` (collective)."""

    SLOT_FLOATS = 8192  # 2 x 4096 channels

    def __init__(self, group, device: torch.device) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        lib = L.load()
        nbytes = int(lib.gdl_p2p_exchange_bytes(self.world, self.SLOT_FLOATS))
        self.buf = symm_mem.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        ptrs = [int(v) for v in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(v == 0 for v in ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer pointers")
        off = self.buf.data_ptr() - ptrs[self.rank]  # the tensor's offset inside its symmetric allocation (same on every rank)
        if off < 0 or off > (1 << 30):
            raise RuntimeError("symmetric memory: the local buffer is not inside the local allocation")
        ptrs = [v + off for v in ptrs]
        self.peer_ptrs = (C.c_void_p * self.world)(*ptrs)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # every rank's buffer is zeroed before anybody signals

    def all_reduce_(self, sums: torch.Tensor) -> torch.Tensor:
        if sums.dtype != torch.float32 or not sums.is_contiguous() or sums.numel() > self.SLOT_FLOATS:
            raise ValueError("P2PExchange: contiguous fp32 vector of at most %d values expected" % self.SLOT_FLOATS)
        _ck(L.load().gdl_p2p_allreduce_sums(L.ptr(sums), sums.numel(), self.peer_ptrs, self.rank, self.world, self.SLOT_FLOATS,
                                            L.ptr(self.counter), L.stream_ptr()))
        return sums


def deterministic() -> bool:
    """ordered (bit-reproducible) reductions are active on the current device (see _lib.stream_ptr)"""
    L.stream_ptr()
    return L.deterministic()


def require_cuda(t: torch.Tensor, what: str) -> None:
    """the hot path has no CPU implementation: refuse host tensors loudly, before any kernel wrapper is reached"""
    if not t.is_cuda:
        raise RuntimeError(f"{what} runs on CUDA (sm_100a) only; there is no CPU fallback")


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
               relu: bool = False, residual: torch.Tensor | None = None, w_ld: int = 0, w_rows: int = 0,
               w_rows_per_img: int = 0, w_mn_major: bool = False, gelu: bool = False,
               oscale: torch.Tensor | None = None, groups: tuple[int, int, int, int] | None = None,
               alg_scale: float = 1.0, bn_sums: torch.Tensor | None = None,
               bn_pivot: torch.Tensor | None = None) -> torch.Tensor:
    """gdl_conv2d_nhwc_fwd. `weight` is the packed [Cout][R][S][Ctot] 16-bit operand (or, with the w_*
    options, a slice of an activation tensor used as the B operand of an attention GEMM).
    alg_scale: ALGORITHMIC / executed FLOPs of this launch (profiler only): < 1 for pixel-packed (block-Toeplitz) and
    channel-padded launches, whose zeros are not work the reference does.
    bn_sums (fp32 [2*Cout]) / bn_pivot (fp32 [Cout] or None): the conv also produces the BatchNorm statistics of its
    rounded output, sum(out - pivot) and sum((out - pivot)^2) per channel — from the epilogue's staged tile where possible."""
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
    d.relu = 2 if gelu else int(relu)
    d.oscale = oscale.data_ptr() if oscale is not None else None
    if residual is not None:
        d.residual = residual.data_ptr()
        d.res_dtype = L.dt_code(residual.dtype)
        d.ldr = residual.stride(2)
    if w_rows_per_img or w_mn_major:  # `weight` is a 2-D view [rows][cols] of an activation tensor
        w_ld, w_rows = weight.stride(0), weight.shape[0]
    d.w_ld, d.w_rows, d.w_rows_per_img, d.w_mn_major = w_ld, w_rows, w_rows_per_img, int(w_mn_major)
    if bn_sums is not None:
        if bn_sums.dtype != torch.float32 or bn_sums.numel() < 2 * cout:
            raise ValueError("bn_sums: fp32 tensor of 2 * Cout elements expected")
        d.bn_sums = bn_sums.data_ptr()
        d.bn_pivot = bn_pivot.data_ptr() if bn_pivot is not None else None
    ng = 1
    if groups is not None:
        # (G, source channel stride, weight stride along its contiguous dim, output channel stride): all heads at once;
        # srcs[0] / weight / out are the views of group 0
        ng, d.g_src_stride, d.g_w_stride, d.g_out_stride = groups
        d.groups = ng
    stats_after = False
    if bn_sums is not None and _PROFILER is not None:
        # per-launch timing of the convolution alone: where the epilogue cannot produce the statistics, run the statistics
        # kernel here (outside the timed bracket) instead of inside the library call — same kernels, same results
        fused = C.c_int(0)
        L.check(L.load().gdl_conv2d_bn_fusable(C.byref(d), C.byref(fused)))
        if not fused.value:
            d.bn_sums, d.bn_pivot, stats_after = None, None, True
    e0 = _PROFILER.begin() if _PROFILER is not None else None
    L.check(L.load().gdl_conv2d_nhwc_fwd(C.byref(d), L.stream_ptr()))
    _count()
    if e0 is not None:
        ctot = sum(t.shape[3] for t in srcs)
        _PROFILER.end("conv_fwd_kernel", 2.0 * alg_scale * ng * n * ho * wo * cout * r * s * ctot, e0,
                      f"N{n} {ho}x{wo} src{[t.shape[3] for t in srcs]} -> {cout} k{r}" + (f" x{ng} groups" if ng > 1 else ""))
    if stats_after:
        bn_stats(out, bn_sums, bn_pivot)
    return out


def conv2d_wgrad(srcs: Sequence[torch.Tensor], dy: torch.Tensor, r: int, s: int, pad_h: int,
                 pad_w: int, dw: torch.Tensor, groups: tuple[int, int, int, int] | None = None,
                 alg_scale: float = 1.0) -> torch.Tensor:
    """gdl_conv2d_nhwc_wgrad: accumulates into fp32 dw.  dw 2-D [Cout][R*S*Ctot] (row stride = stride(0)), or
    3-D [N][Cout][Ctot] = batched: one independent product per image (attention dV = P^T dO, dK = dS^T q).
    groups = (G, source channel stride, dy channel stride, dw column stride): G such products per image in one launch (all
    heads); srcs[0] / dy / dw are the views of group 0."""
    d = L.ConvWgrad()
    _fill_srcs(d, srcs)
    d.Cout = dy.shape[3]
    d.R, d.S, d.pad_h, d.pad_w = r, s, pad_h, pad_w
    d.dy = dy.data_ptr()
    d.ld_dy = dy.stride(2)
    d.dtype = L.dt_code(srcs[0].dtype)
    if dw.dtype != torch.float32 or dw.stride(-1) != 1:
        raise ValueError("dw must be fp32 with a contiguous last dim")
    d.dw = dw.data_ptr()
    if dw.dim() == 3:
        d.batched, d.dw_img_stride, d.dw_ld = 1, dw.stride(0), dw.stride(1)
    else:
        d.batched, d.dw_img_stride, d.dw_ld = 0, 0, dw.stride(0)
    if groups is not None and groups[0] > 1:
        d.groups, d.g_src_stride, d.g_dy_stride, d.g_dw_stride = groups
    e0 = _PROFILER.begin() if _PROFILER is not None else None
    L.check(L.load().gdl_conv2d_nhwc_wgrad(C.byref(d), L.stream_ptr()))
    _count()
    if e0 is not None:
        ctot = sum(t.shape[3] for t in srcs)
        _PROFILER.end("conv_wgrad_kernel", 2.0 * alg_scale * (d.groups or 1) * dy.shape[0] * dy.shape[1] * dy.shape[2] * dy.shape[3] * r * s * ctot, e0,
                      f"N{dy.shape[0]} {dy.shape[1]}x{dy.shape[2]} src{[t.shape[3] for t in srcs]} -> {dy.shape[3]} k{r}")
    return dw


# ---------------------------------------------------------------------------------------------
# fp32-accurate tensor-core convolution: two bf16 planes per operand (north_star tolerance 1e-5 for fp32)
#   x = xh + xl,  w = wh + wl  with  xh = bf16(x), xl = bf16(x - xh)  (|xl| <= 2^-9 |x|, |x - xh - xl| <= 2^-18 |x|)
#   x * w  =  xh*wh + xl*wh + xh*wl + xl*wl      — four bf16 MMAs with fp32 accumulation in TMEM.
# The launches chain through the kernel's fp32 residual input (out += ...), so nothing but the final fp32 tensor is
# written.  4x the tensor work of the 16-bit path: a verification / high-accuracy mode, not the training default.
# (Measured on B200, max-norm error against torch's fp32 convolution: 4e-6 .. 7e-6 forward / data gradient; the weight
# gradient needed the fourth, lo x lo, product to stay under 1e-5 — with three it sat at 1.00e-5 .. 1.04e-5.)
# ---------------------------------------------------------------------------------------------
def split_bf16(x32: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    hi = x32.to(torch.bfloat16)
    lo = (x32 - hi.float()).to(torch.bfloat16)
    return hi, lo


def conv2d_fwd_bf16x2(srcs32: Sequence[torch.Tensor], w_oihw32: torch.Tensor, pad_h: int, pad_w: int, *,
                      bias: torch.Tensor | None = None, mode: int = 0) -> torch.Tensor:
    """fp32 NHWC sources (virtual concat) x fp32 OIHW weight -> fp32 NHWC output, stride 1, accurate to < 1e-5 (max norm).
    mode 1 = the data gradient: `srcs32` is dY and the weight is applied transposed / tap-flipped (pad = R-1-pad)."""
    k, c, r, s_ = w_oihw32.shape
    cout = k if mode == 0 else c
    wh, wl = split_bf16(w_oihw32.contiguous())
    wph = pack_conv_weight(wh.float(), torch.bfloat16, mode)
    wpl = pack_conv_weight(wl.float(), torch.bfloat16, mode)
    hi, lo = zip(*[split_bf16(t) for t in srcs32])
    out = conv2d_fwd(list(lo), wpl, cout, r, s_, pad_h, pad_w, out_dtype=torch.float32)  # smallest terms first
    conv2d_fwd(list(lo), wph, cout, r, s_, pad_h, pad_w, out=out, residual=out)
    conv2d_fwd(list(hi), wpl, cout, r, s_, pad_h, pad_w, out=out, residual=out)
    conv2d_fwd(list(hi), wph, cout, r, s_, pad_h, pad_w, out=out, residual=out, bias=bias)
    return out


def conv2d_wgrad_bf16x2(srcs32: Sequence[torch.Tensor], dy32: torch.Tensor, r: int, s: int, pad_h: int, pad_w: int) -> torch.Tensor:
    """fp32 weight gradient [Cout][R*S*Ctot] of fp32 operands through the four bf16 cross products of their two planes"""
    hi, lo = zip(*[split_bf16(t) for t in srcs32])
    dyh, dyl = split_bf16(dy32)
    ctot = sum(t.shape[3] for t in srcs32)
    dw = torch.zeros((dy32.shape[3], r * s * ctot), dtype=torch.float32, device=dy32.device)
    conv2d_wgrad(list(lo), dyl, r, s, pad_h, pad_w, dw)
    conv2d_wgrad(list(lo), dyh, r, s, pad_h, pad_w, dw)
    conv2d_wgrad(list(hi), dyl, r, s, pad_h, pad_w, dw)
    conv2d_wgrad(list(hi), dyh, r, s, pad_h, pad_w, dw)
    return dw


def pack_conv_weight(w_oihw: torch.Tensor, dtype: torch.dtype, mode: int = 0, ld: int = 0,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 OIHW -> 16-bit operand. mode 0: [Cout][(r,s,c)]; 1: [Cin][(R-1-r,S-1-s,k)] (dgrad);
    2: [(r,s,c)][Cout] (dgrad of an im2col'd conv). `ld` pads the row stride with zeros."""
    k, c, r, s = w_oihw.shape
    if w_oihw.dtype != torch.float32 or not w_oihw.is_contiguous():
        raise ValueError("weight must be contiguous fp32 OIHW")
    rows = k if mode == 0 else (c if mode == 1 else r * s * c)
    cols = r * s * c if mode == 0 else (r * s * k if mode == 1 else k)
    ld = ld or cols
    if out is None:
        out = torch.empty((rows, ld), dtype=dtype, device=w_oihw.device)
    _ck(L.load().gdl_pack_conv_weight(L.ptr(w_oihw), L.ptr(out), k, c, r, s, mode, ld,
                                          L.dt_code(dtype), L.stream_ptr()))
    return out


def unpack_conv_wgrad(dw: torch.Tensor, out_oihw: torch.Tensor, src_ld: int = 0,
                      accumulate: bool = False) -> torch.Tensor:
    k, c, r, s = out_oihw.shape
    _ck(L.load().gdl_unpack_conv_wgrad(L.ptr(dw), L.ptr(out_oihw), k, c, r, s, src_ld,
                                           int(accumulate), L.stream_ptr()))
    return out_oihw


REPACK_CHUNK = 4096  # kRepackChunk of igemm_conv.cu


def repack_weights(entries: list[tuple[torch.Tensor, torch.Tensor, int, int, int, int, int, int]], plan=None):
    """Refresh every packed 16-bit operand in `entries` = [(fp32 OIHW master, packed out, Cout, Cin, R, S, mode, dst_ld)] from
    its master in ONE launch (gdl_repack_weights).  Returns the device-side plan; pass it back while `entries` is unchanged."""
    if not entries:
        return plan
    if plan is None:
        tab = (L.Repack * len(entries))()
        chunk0 = [0]
        for i, (w, out, co, ci, r, s_, mode, ld) in enumerate(entries):
            tab[i].src, tab[i].dst = w.data_ptr(), out.data_ptr()
            tab[i].Cout, tab[i].Cin, tab[i].R, tab[i].S, tab[i].mode, tab[i].dst_ld = co, ci, r, s_, mode, ld
            rows = co if mode == 0 else (ci if mode == 1 else r * s_ * ci)
            chunk0.append(chunk0[-1] + (rows * ld + REPACK_CHUNK - 1) // REPACK_CHUNK)
        dev = entries[0][1].device
        raw = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(dev)
        plan = (raw, torch.tensor(chunk0, dtype=torch.int32, device=dev), len(entries), chunk0[-1],
                L.dt_code(entries[0][1].dtype))
    raw, chunk0_dev, n, total, dtc = plan
    _ck(L.load().gdl_repack_weights(L.ptr(raw), L.ptr(chunk0_dev), n, total, dtc, L.stream_ptr()))
    return plan


def widen_conv_weight(wp: torch.Tensor, co: int, ci: int, r: int, f: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """packed 16-bit [Co][R][3][Ci] -> block-Toeplitz [f*Co][R][3][f*Ci]: the same conv on (N,H,W/f,f*Ci) pixels"""
    if wp.shape != (co, r * 3 * ci) or not wp.is_contiguous():
        raise ValueError("widen_conv_weight: dense packed [Co][R*3*Ci] operand expected")
    if out is None:
        out = torch.empty((f * co, r * 3 * f * ci), dtype=wp.dtype, device=wp.device)
    _ck(L.load().gdl_widen_conv_weight(L.ptr(wp), L.ptr(out), co, ci, r, f, L.dt_code(wp.dtype), L.stream_ptr()))
    return out


def fold_widened_wgrad(dw: torch.Tensor, out_oihw: torch.Tensor, f: int, accumulate: bool = False) -> torch.Tensor:
    """fp32 gradient of widened weights [f*Co][R*3*f*Ci] -> fp32 OIHW gradient of the real Rx3 conv"""
    co, ci, r, s = out_oihw.shape
    src_co = dw.shape[0] // f  # >= co when the conv's output channels were zero-padded
    if s != 3 or dw.shape[0] != f * src_co or src_co < co:
        raise ValueError("fold_widened_wgrad: shape mismatch")
    _ck(L.load().gdl_fold_widened_wgrad(L.ptr(dw), dw.stride(0), src_co, L.ptr(out_oihw), co, ci, r, f, int(accumulate),
                                        L.stream_ptr()))
    return out_oihw


# ---------------------------------------------------------------------------------------------
# HBM-bound kernels (elementwise.cu)
# ---------------------------------------------------------------------------------------------
_IN_KIND = {(torch.uint8, False): 0, (torch.float32, False): 1, (torch.float32, True): 2, (torch.uint8, True): 3}


def normalize_to_nhwc(x: torch.Tensor, chw: bool, out_dtype: torch.dtype, ld: int,
                      mean: torch.Tensor | None = None, std: torch.Tensor | None = None,
                      image_max: float = 0.0) -> torch.Tensor:
    """uint8/f32 image batch (NHWC if chw=False else NCHW) -> 16-bit NHWC (N,H,W,ld), channels >= C zero.
    y = ((x / image_max) - mean) / std   (utils/tensors.py:10-35); image_max=0 and mean=None skip a step."""
    if not x.is_contiguous():
        raise ValueError("normalize: contiguous input expected")
    if chw:
        n, c, h, w = x.shape
    else:
        n, h, w, c = x.shape
    out = torch.empty((n, h, w, ld), dtype=out_dtype, device=x.device)
    _ck(L.load().gdl_normalize_to_nhwc(L.ptr(x), _IN_KIND[(x.dtype, chw)], L.ptr(out), L.dt_code(out_dtype),
                                           n, h, w, c, ld, L.ptr(mean), L.ptr(std), float(image_max),
                                           L.stream_ptr()))
    return out


AUG_IDENTITY, AUG_HFLIP, AUG_VFLIP, AUG_ROT90, AUG_CROP = range(5)


def augment_normalize(x: torch.Tensor, chw: bool, mask: torch.Tensor | None, params: torch.Tensor,
                      out_dtype: torch.dtype, ld: int = 0, mean: torch.Tensor | None = None,
                      std: torch.Tensor | None = None, image_max: float = 0.0):
    """Batch augmentation fused with the patch normalisation (gdl_augment_normalize): x uint8/f32 image batch (NHWC if
    chw=False else NCHW), mask (N,H,W) int64/uint8 or None, params int32 (N,6) = {op, k, y0, x0, ch, cw} on the device.
    Returns (image, mask'): image = 16-bit NHWC (N,H,W,ld) for a 16-bit out_dtype, f32 NCHW for torch.float32."""
    if not x.is_contiguous() or (mask is not None and not mask.is_contiguous()):
        raise ValueError("augment: contiguous inputs expected")
    if chw:
        n, c, h, w = x.shape
    else:
        n, h, w, c = x.shape
    if params.dtype != torch.int32 or tuple(params.shape) != (n, 6) or not params.is_contiguous():
        raise ValueError(f"augment: params must be a contiguous int32 ({n}, 6) tensor")
    if mask is not None and tuple(mask.shape) != (n, h, w):
        raise ValueError(f"augment: mask shape {tuple(mask.shape)} does not match the image batch {(n, h, w)}")
    if out_dtype == torch.float32:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    else:
        ld = ld or (c + 7) // 8 * 8
        out = torch.empty((n, h, w, ld), dtype=out_dtype, device=x.device)
    mask_out = torch.empty_like(mask) if mask is not None else None
    _ck(L.load().gdl_augment_normalize(L.ptr(x), _IN_KIND[(x.dtype, chw)], L.ptr(mask),
                                           _target_kind(mask) if mask is not None else 0, L.ptr(params), L.ptr(out),
                                           L.dt_code(out_dtype), L.ptr(mask_out), n, h, w, c, ld, L.ptr(mean), L.ptr(std),
                                           float(image_max), L.stream_ptr()))
    return out, mask_out


def im2col(x: torch.Tensor, c: int, r: int, s: int, stride: int, pad: int, kpad: int) -> torch.Tensor:
    n, h, w, _ = x.shape
    ho, wo = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1
    col = torch.empty((n, ho, wo, kpad), dtype=x.dtype, device=x.device)
    _ck(L.load().gdl_im2col_nhwc(L.ptr(x), L.ptr(col), L.dt_code(x.dtype), n, h, w, c, x.stride(2), r, s,
                                     stride, pad, kpad, L.stream_ptr()))
    return col


def col2im(dcol: torch.Tensor, n: int, h: int, w: int, c: int, r: int, s: int, stride: int, pad: int) -> torch.Tensor:
    dx = torch.empty((n, h, w, c), dtype=dcol.dtype, device=dcol.device)
    _ck(L.load().gdl_col2im_nhwc(L.ptr(dcol), L.ptr(dx), L.dt_code(dcol.dtype), n, h, w, c, c, r, s, stride,
                                     pad, dcol.stride(2), L.stream_ptr()))
    return dx


def _rows(x: torch.Tensor) -> int:
    return x.shape[0] * x.shape[1] * x.shape[2]


def bn_stats(x: torch.Tensor, sums: torch.Tensor, pivot: torch.Tensor | None = None) -> torch.Tensor:
    """sums[0:C] = sum(x - p), sums[C:2C] = sum((x-p)^2); p = per-channel pivot (fp32 [C]) or 0."""
    _ck(L.load().gdl_bn_stats(L.ptr(x), L.dt_code(x.dtype), _rows(x), x.shape[3], x.stride(2), L.ptr(sums),
                                  L.ptr(pivot), L.stream_ptr()))
    return sums


def bn_finalize(pivot, sums, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, save_mean,
                save_invstd) -> None:
    _ck(L.load().gdl_bn_finalize(L.ptr(pivot), L.ptr(sums), int(count), scale.numel(), L.ptr(gamma),
                                     L.ptr(beta), float(eps), float(momentum), L.ptr(running_mean),
                                     L.ptr(running_var), L.ptr(scale), L.ptr(shift), L.ptr(save_mean),
                                     L.ptr(save_invstd), L.stream_ptr()))


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps, scale, shift) -> None:
    _ck(L.load().gdl_bn_eval_coeffs(running_mean.numel(), L.ptr(gamma), L.ptr(beta), L.ptr(running_mean),
                                        L.ptr(running_var), float(eps), L.ptr(scale), L.ptr(shift), L.stream_ptr()))


def bn_apply(x, scale, shift, *, res=None, rscale=None, rshift=None, relu=True, y=None, y_up=None) -> None:
    n, h, w, c = x.shape
    _ck(L.load().gdl_bn_apply(L.ptr(x), x.stride(2), L.ptr(scale), L.ptr(shift), L.ptr(res),
                                  res.stride(2) if res is not None else 0, L.ptr(rscale), L.ptr(rshift), int(relu),
                                  L.ptr(y), y.stride(2) if y is not None else 0, L.ptr(y_up),
                                  y_up.stride(2) if y_up is not None else 0, L.dt_code(x.dtype), n, h, w, c,
                                  L.stream_ptr()))


def grad_gather(srcs, shape, dtype, *, y=None, x=None, mean=None, invstd=None, g=None, sums=None) -> None:
    """g = (sum of gradient sources) [* (y > 0)]; srcs = [(tensor, mode)], mode 1 = 2x2 sum of a 2H x 2W tensor.
    With sums/x/mean/invstd also reduces the BatchNorm-backward sums in the same pass."""
    n, h, w, c = shape
    k = len(srcs)
    ptrs = (C.c_void_p * k)(*[t.data_ptr() for t, _ in srcs])
    lds = (C.c_int * k)(*[t.stride(2) for t, _ in srcs])
    modes = (C.c_int * k)(*[m for _, m in srcs])
    _ck(L.load().gdl_grad_gather(k, ptrs, lds, modes, L.ptr(y), y.stride(2) if y is not None else 0, L.ptr(x),
                                     x.stride(2) if x is not None else 0, L.ptr(mean), L.ptr(invstd), L.ptr(g),
                                     g.stride(2) if g is not None else 0, L.ptr(sums), L.dt_code(dtype), n, h, w, c,
                                     L.stream_ptr()))


def bn_bwd_apply(g, x, mean, invstd, gamma, sums, dx, dgamma, dbeta, accumulate: bool, count: int = 0) -> None:
    """count = number of elements per channel the (possibly all-reduced) sums cover (0 = local rows)."""
    _ck(L.load().gdl_bn_bwd_apply(L.ptr(g), g.stride(2), L.ptr(x), x.stride(2), L.ptr(mean), L.ptr(invstd),
                                      L.ptr(gamma), L.ptr(sums), L.ptr(dx), dx.stride(2), L.ptr(dgamma),
                                      L.ptr(dbeta), int(accumulate), L.dt_code(x.dtype), _rows(x),
                                      int(count) or _rows(x), x.shape[3], L.stream_ptr()))


def bn_param_grads(sums, dgamma, dbeta, accumulate: bool = False) -> None:
    """dgamma = sums[C:2C] (sum g*xhat), dbeta = sums[0:C] (sum g)."""
    _ck(L.load().gdl_bn_param_grads(L.ptr(sums), sums.numel() // 2, L.ptr(dgamma), L.ptr(dbeta),
                                        int(accumulate), L.stream_ptr()))


def maxpool3x3s2_fwd(x: torch.Tensor, want_idx: bool):
    n, h, w, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if want_idx else None
    _ck(L.load().gdl_maxpool3x3s2_fwd(L.ptr(x), x.stride(2), L.ptr(y), c, L.ptr(idx), L.dt_code(x.dtype), n, h, w,
                                          c, L.stream_ptr()))
    return y, idx


def maxpool3x3s2_bwd(dy: torch.Tensor, idx: torch.Tensor, h: int, w: int) -> torch.Tensor:
    n, _, _, c = dy.shape
    dx = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device)
    _ck(L.load().gdl_maxpool3x3s2_bwd(L.ptr(dy), dy.stride(2), L.ptr(idx), L.ptr(dx), c, L.dt_code(dy.dtype), n,
                                          h, w, c, L.stream_ptr()))
    return dx


# ---------------------------------------------------------------------------------------------
# losses / optimizer (loss_optim.cu)
# ---------------------------------------------------------------------------------------------
class LossSpec:
    """weights of the CE and Dice terms + their options (see csrc/loss_optim.cu)."""

    def __init__(self, w_ce=1.0, w_dice=0.0, label_smoothing=0.0, ce_mean_over_all=False, ignore_index=None,
                 dice_smooth=0.0, dice_eps=1e-7):
        self.w_ce, self.w_dice = float(w_ce), float(w_dice)
        self.label_smoothing = float(label_smoothing)
        self.ce_mean_over_all = bool(ce_mean_over_all)
        self.ignore_index = ignore_index
        self.dice_smooth, self.dice_eps = float(dice_smooth), float(dice_eps)

    def args(self):
        ign = self.ignore_index
        return (int(ign) if ign is not None else 0, int(ign is not None), self.w_ce, self.w_dice,
                self.label_smoothing, int(self.ce_mean_over_all), self.dice_smooth, self.dice_eps)


def _target_kind(t: torch.Tensor) -> int:
    if t.dtype == torch.int64:
        return 0
    if t.dtype == torch.uint8:
        return 1
    raise ValueError(f"target dtype must be int64 or uint8, got {t.dtype}")


def seg_loss_fwd(logits: torch.Tensor, target: torch.Tensor, spec: LossSpec):
    """logits fp32 (N,H,W,K) NHWC (pixel stride = stride(2)); returns (coeff, stats); coeff[0] is the loss."""
    k = logits.shape[3]
    m = _rows(logits)
    stats = torch.empty(4 + 3 * k, dtype=torch.float32, device=logits.device)
    coeff = torch.empty(2 + 2 * k, dtype=torch.float32, device=logits.device)
    _ck(L.load().gdl_seg_loss_fwd(L.ptr(logits), logits.stride(2), L.ptr(target), _target_kind(target), m, k,
                                      *spec.args(), L.ptr(stats), L.ptr(coeff), L.stream_ptr()))
    return coeff, stats


def seg_loss_bwd(logits, target, spec: LossSpec, coeff, grad_scale, dlogits) -> None:
    k = logits.shape[3]
    _ck(L.load().gdl_seg_loss_bwd(L.ptr(logits), logits.stride(2), L.ptr(target), _target_kind(target),
                                      _rows(logits), k, *spec.args(), L.ptr(coeff), L.ptr(grad_scale), L.ptr(dlogits),
                                      dlogits.stride(2), L.dt_code(dlogits.dtype), L.stream_ptr()))


def upsample_ce_fwd(logits_lr: torch.Tensor, target: torch.Tensor, spec: LossSpec):
    """Loss of bilinear(logits_lr -> target's H x W) against target without materialising the upsampled logits.
    logits_lr fp32 (N,h,w,K) NHWC; target (N,H,W) int64 / uint8.  Returns (coeff, stats) as seg_loss_fwd."""
    n, h, w, k = logits_lr.shape
    hh, ww = target.shape[1:3]
    stats = torch.empty(4 + 3 * k, dtype=torch.float32, device=logits_lr.device)
    coeff = torch.empty(2 + 2 * k, dtype=torch.float32, device=logits_lr.device)
    _ck(L.load().gdl_upsample_ce_fwd(L.ptr(logits_lr), logits_lr.stride(2), n, h, w, hh, ww, L.ptr(target),
                                     _target_kind(target), k, *spec.args(), L.ptr(stats), L.ptr(coeff), L.stream_ptr()))
    return coeff, stats


def upsample_ce_bwd(logits_lr, target, spec: LossSpec, coeff, grad_scale, dlogits_lr) -> None:
    """dlogits_lr (N,h,w,>=K) fp32 or 16-bit: d(loss)/d(logits_lr) [* grad_scale]; channels >= K are not written."""
    n, h, w, k = logits_lr.shape
    hh, ww = target.shape[1:3]
    _ck(L.load().gdl_upsample_ce_bwd(L.ptr(logits_lr), logits_lr.stride(2), n, h, w, hh, ww, L.ptr(target),
                                     _target_kind(target), k, *spec.args(), L.ptr(coeff), L.ptr(grad_scale),
                                     L.ptr(dlogits_lr), dlogits_lr.stride(2), L.dt_code(dlogits_lr.dtype), L.stream_ptr()))


def upsample_argmax(logits_lr: torch.Tensor, hh: int, ww: int, threshold: float = 0.5) -> torch.Tensor:
    """argmax over classes (K == 1: sigmoid > threshold) of bilinear(logits_lr -> hh x ww): (N,hh,ww) int64."""
    n, h, w, k = logits_lr.shape
    out = torch.empty((n, hh, ww), dtype=torch.int64, device=logits_lr.device)
    _ck(L.load().gdl_upsample_argmax(L.ptr(logits_lr), logits_lr.stride(2), n, h, w, hh, ww, k, float(threshold),
                                     L.ptr(out), L.stream_ptr()))
    return out


def argmax_classes(logits: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    n, h, w, k = logits.shape
    out = torch.empty((n, h, w), dtype=torch.int64, device=logits.device)
    _ck(L.load().gdl_argmax_classes(L.ptr(logits), logits.stride(2), n * h * w, k, float(threshold), L.ptr(out),
                                        L.stream_ptr()))
    return out


def argmax_confusion(logits: torch.Tensor, target: torch.Tensor | None, threshold: float = 0.5,
                     ignore_index: int | None = None, want_classes: bool = True):
    """logits fp32 (N,H,W,K) -> (classes int64 (N,H,W) | None, conf int64 (N,Kc,Kc) | None): the eval post-processing
    (softmax.argmax / sigmoid > threshold) fused with the per-sample confusion counts conf[n][target][prediction]."""
    n, h, w, k = logits.shape
    kc = 2 if k == 1 else k
    classes = torch.empty((n, h, w), dtype=torch.int64, device=logits.device) if want_classes else None
    conf = None
    if target is not None:
        if tuple(target.shape) != (n, h, w) or not target.is_contiguous():
            raise ValueError(f"argmax_confusion: target must be a contiguous {(n, h, w)} tensor")
        conf = torch.zeros((n, kc, kc), dtype=torch.int64, device=logits.device)
    _ck(L.load().gdl_argmax_confusion(L.ptr(logits), logits.stride(2), n, h * w, k, float(threshold), L.ptr(target),
                                          _target_kind(target) if target is not None else 0,
                                          int(ignore_index) if ignore_index is not None else 0,
                                          int(ignore_index is not None), L.ptr(classes), L.ptr(conf), L.stream_ptr()))
    return classes, conf


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None) -> None:
    _ck(L.load().gdl_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), float(lr), float(beta1),
                                   float(beta2), float(eps), float(weight_decay), int(step), L.ptr(grad_scale),
                                   L.stream_ptr()))


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, weight_decay, state, grad_scale=None, lr_scale=None) -> None:
    """Adam with the step counter / bias corrections kept on the device (`state`: 3 floats); lr_scale: optional device
    scalar multiplying lr (a scheduler's factor: read at run time, so a captured CUDA graph follows it)."""
    _ck(L.load().gdl_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), float(lr), float(beta1),
                                   float(beta2), float(eps), float(weight_decay), L.ptr(state), L.ptr(grad_scale),
                                   L.ptr(lr_scale), L.stream_ptr()))


def grad_clip_coef(g, max_norm, scratch, scale) -> None:
    _ck(L.load().gdl_grad_clip_coef(L.ptr(g), g.numel(), float(max_norm), L.ptr(scratch), L.ptr(scale),
                                        L.stream_ptr()))


# ---------------------------------------------------------------------------------------------
# MixTransformer / SegFormer kernels (transformer.cu).  Token tensors are (..., C) with a uniform row
# stride stride(-2); M = number of rows.
# ---------------------------------------------------------------------------------------------
def _rows2(t: torch.Tensor) -> int:
    return t.numel() // t.shape[-1]


def layernorm_fwd(x, gamma, beta, eps, out_dtype, want_stats: bool = True):
    m, c = _rows2(x), x.shape[-1]
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    stats = torch.empty((2, m), dtype=torch.float32, device=x.device) if want_stats else None
    _ck(L.load().gdl_layernorm_fwd(L.ptr(x), L.dt_code(x.dtype), x.stride(-2), L.ptr(gamma), L.ptr(beta), float(eps),
                                   L.ptr(y), L.dt_code(out_dtype), c, L.ptr(stats[0]) if want_stats else None,
                                   L.ptr(stats[1]) if want_stats else None, m, c, L.stream_ptr()))
    return y, stats


def layernorm_bwd(g, x, stats, gamma, *, add=None, want32: bool = True, dtype16=None, pgrads=None):
    """returns (dx32 | None, dx16 | None); pgrads [2][C] (dgamma, dbeta) is accumulated into."""
    m, c = _rows2(x), x.shape[-1]
    dx32 = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want32 else None
    dx16 = torch.empty(x.shape, dtype=dtype16, device=x.device) if dtype16 is not None else None
    _ck(L.load().gdl_layernorm_bwd(L.ptr(g), L.dt_code(g.dtype), g.stride(-2), L.ptr(x), L.dt_code(x.dtype),
                                   x.stride(-2), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(gamma), L.ptr(add),
                                   add.stride(-2) if add is not None else 0, L.ptr(dx32), c, L.ptr(dx16),
                                   L.dt_code(dtype16) if dtype16 is not None else 0, c, L.ptr(pgrads), m, c,
                                   L.stream_ptr()))
    return dx32, dx16


def softmax_fwd(s, scale, length, p=None):
    """s: (..., Lpad) rows; p = softmax(scale * s[..., :length]) with zeros in the pad columns."""
    lpad = s.shape[-1]
    if p is None:
        p = torch.empty_like(s)
    _ck(L.load().gdl_softmax_fwd(L.ptr(s), s.stride(-2), float(scale), L.ptr(p), p.stride(-2), L.dt_code(s.dtype),
                                 _rows2(s), length, lpad, L.stream_ptr()))
    return p


def sra_attention_supported(n: int, nk: int, d: int, save_p: bool = True) -> bool:
    """shapes gdl_sra_attention_fwd / _bwd cover.  Training (probabilities saved for the backward): all keys of a head in one TMEM
    tile, whole 64-key store boxes — every MiT stage of a 256x256 or 512x512 tile.  Inference: any token counts (streamed keys)."""
    if d != 64 or n <= 0:
        return False
    return not save_p or (nk % 64 == 0 and 0 < nk <= 256)


def sra_attention_fwd(q, kv2, heads, nk, scale, save_p=True):
    """Fused attention forward.  q: (B, N, c) 16-bit tokens; kv2: (B*nk, 2c) reduced tokens (K | V).
    Returns (o (B, N, c), p (B, N, heads*nk) or None)."""
    b, n, c = q.shape
    if q.stride(2) != 1 or kv2.stride(1) != 1 or q.stride(0) != n * q.stride(1):
        raise ValueError("sra_attention_fwd: token rows with a contiguous channel dim expected")
    o = torch.empty((b, n, c), dtype=q.dtype, device=q.device)
    p = torch.empty((b, n, heads * nk), dtype=q.dtype, device=q.device) if save_p else None
    _ck(L.load().gdl_sra_attention_fwd(L.ptr(q), q.stride(1), L.ptr(kv2), kv2.stride(0), L.ptr(o), o.stride(1), L.ptr(p),
                                       p.stride(1) if p is not None else 0, b, n, heads, nk, c, float(scale),
                                       L.dt_code(q.dtype), L.stream_ptr()))
    return o, p


def sra_attention_bwd(do, kv2, p, heads, nk, scale):
    """Fused dP -> dS -> dQ.  do: (B, N, c) gradient of the attention output; kv2: (B*nk, 2c); p: (B, N, heads*nk) saved
    probabilities.  Returns (dq (B, N, c), ds (B, N, heads*nk))."""
    b, n, c = do.shape
    if do.stride(2) != 1 or p.stride(2) != 1 or kv2.stride(1) != 1 or do.stride(0) != n * do.stride(1) or p.stride(0) != n * p.stride(1):
        raise ValueError("sra_attention_bwd: token rows with a contiguous channel dim expected")
    dq = torch.empty((b, n, c), dtype=do.dtype, device=do.device)
    ds = torch.empty((b, n, heads * nk), dtype=do.dtype, device=do.device)
    _ck(L.load().gdl_sra_attention_bwd(L.ptr(do), do.stride(1), L.ptr(kv2), kv2.stride(0), L.ptr(p), p.stride(1), L.ptr(dq), dq.stride(1),
                                       L.ptr(ds), ds.stride(1), b, n, heads, nk, c, float(scale), L.dt_code(do.dtype), L.stream_ptr()))
    return dq, ds


def mha_flash_fwd(qkv, b, n, heads, scale):
    """Self-attention forward of a fused projection qkv (b*n, 3c) = [q | k | v], head dim 64, any token count: one kernel, keys
    streamed in blocks with the online softmax.  Returns o (b*n, c)."""
    c = qkv.shape[1] // 3
    if qkv.dim() != 2 or qkv.stride(1) != 1 or qkv.shape[0] != b * n or c != 64 * heads:
        raise ValueError("mha_flash_fwd: (b*n, 3*64*heads) projection with a contiguous channel dim expected")
    o = torch.empty((b * n, c), dtype=qkv.dtype, device=qkv.device)
    _ck(L.load().gdl_mha_flash_fwd(L.ptr(qkv), qkv.stride(0), L.ptr(qkv[:, c:]), qkv.stride(0), L.ptr(qkv[:, 2 * c:]), qkv.stride(0),
                                   L.ptr(o), o.stride(0), b, n, heads, float(scale), L.dt_code(qkv.dtype), L.stream_ptr()))
    return o


def softmax_bwd(p, dp, scale, length, ds=None):
    lpad = p.shape[-1]
    if ds is None:
        ds = torch.empty_like(p)
    _ck(L.load().gdl_softmax_bwd(L.ptr(p), p.stride(-2), L.ptr(dp), dp.stride(-2), float(scale), L.ptr(ds),
                                 ds.stride(-2), L.dt_code(p.dtype), _rows2(p), length, lpad, L.stream_ptr()))
    return ds


def dwconv3x3_gelu_fwd(x, w, bias):
    """x (N,H,W,C) 16-bit; w fp32 [C][3][3] (nn.Conv2d(groups=C).weight squeezed); returns (y, pre)."""
    n, h, wd, c = x.shape
    pre = torch.empty((n, h, wd, c), dtype=x.dtype, device=x.device)
    y = torch.empty_like(pre)
    _ck(L.load().gdl_dwconv3x3_gelu_fwd(L.ptr(x), x.stride(2), L.ptr(w), L.ptr(bias), L.ptr(pre), L.ptr(y),
                                        L.dt_code(x.dtype), n, h, wd, c, L.stream_ptr()))
    return y, pre


def dwconv3x3_gelu_bwd(dy, pre, x, w, pgrads):
    """returns dx; pgrads fp32 [C][10] (9 taps + bias) accumulated into."""
    n, h, wd, c = x.shape
    scratch = torch.empty((n, h, wd, c), dtype=x.dtype, device=x.device)
    dx = torch.empty((n, h, wd, c), dtype=x.dtype, device=x.device)
    _ck(L.load().gdl_dwconv3x3_gelu_bwd(L.ptr(dy), L.ptr(pre), L.ptr(x), x.stride(2), L.ptr(w), L.ptr(scratch),
                                        L.ptr(dx), c, L.ptr(pgrads), L.dt_code(x.dtype), n, h, wd, c, L.stream_ptr()), )
    return dx


def bilinear_fwd(x, ho, wo, out=None):
    n, hi, wi, c = x.shape
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
    _ck(L.load().gdl_bilinear_fwd(L.ptr(x), x.stride(2), L.ptr(out), out.stride(2), L.dt_code(x.dtype), n, hi, wi, ho,
                                  wo, c, L.stream_ptr()))
    return out


def bilinear_sum_fwd(base, srcs, out=None):
    """base (N,Ho,Wo,C) + sum of the (N,h_i,w_i,C) maps in `srcs` (<= 3), each resized bilinearly (align_corners=False) to
    (Ho,Wo): fp32 sum in the given order, one rounding (gdl_bilinear_sum_fwd).  16-bit NHWC, channel slices allowed."""
    import ctypes as C
    n, ho, wo, c = base.shape
    require_cuda(base, "bilinear_sum_fwd")
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=base.dtype, device=base.device)
    arr = (L.Lowres * max(1, len(srcs)))()
    for i, t in enumerate(srcs):
        if t.shape[0] != n or t.shape[3] != c or t.dtype != base.dtype:
            raise ValueError("bilinear_sum_fwd: sources must share batch, channels and dtype with the base")
        arr[i].ptr, arr[i].H, arr[i].W, arr[i].ld = t.data_ptr(), t.shape[1], t.shape[2], t.stride(2)
    _ck(L.load().gdl_bilinear_sum_fwd(L.ptr(base), base.stride(2), len(srcs), C.cast(arr, C.c_void_p), L.ptr(out),
                                      out.stride(2), L.dt_code(base.dtype), n, ho, wo, c, L.stream_ptr()))
    return out


def bilinear_bwd(dy, hi, wi):
    n, ho, wo, c = dy.shape
    dx = torch.empty((n, hi, wi, c), dtype=dy.dtype, device=dy.device)
    _ck(L.load().gdl_bilinear_bwd(L.ptr(dy), dy.stride(2), L.ptr(dx), c, L.dt_code(dy.dtype), n, hi, wi, ho, wo, c,
                                  L.stream_ptr()))
    return dx


def cast_f32(x, dtype):
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    _ck(L.load().gdl_cast_f32(L.ptr(x), L.ptr(y), L.dt_code(dtype), x.numel(), L.stream_ptr()))
    return y


def adaptive_avgpool_fwd(x, s):
    n, h, w, c = x.shape
    y = torch.empty((n, s, s, c), dtype=x.dtype, device=x.device)
    _ck(L.load().gdl_adaptive_avgpool_fwd(L.ptr(x), x.stride(2), L.ptr(y), L.dt_code(x.dtype), n, h, w, c, s, L.stream_ptr()))
    return y


def adaptive_avgpool_bwd(dy, h, w):
    n, s, _, c = dy.shape
    if not dy.is_contiguous():
        raise ValueError("adaptive_avgpool_bwd: dense dy expected")
    dx = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device)
    _ck(L.load().gdl_adaptive_avgpool_bwd(L.ptr(dy), L.ptr(dx), L.dt_code(dy.dtype), n, h, w, c, s, L.stream_ptr()))
    return dx


def add_nhwc(a, b):
    n, h, w, c = a.shape
    y = torch.empty((n, h, w, c), dtype=a.dtype, device=a.device)
    _ck(L.load().gdl_add_nhwc(L.ptr(a), a.stride(2), L.ptr(b), b.stride(2), L.ptr(y), c, L.dt_code(a.dtype), n * h * w, c,
                              L.stream_ptr()))
    return y


def vit_assemble_tokens(patch, pos, cls):
    """patch (B,P,C) 16-bit/fp32, pos fp32 (P+1,C), cls fp32 (C,) -> fp32 tokens (B,P+1,C) = [cls ; patch + pos[1:]]"""
    b, p_, c = patch.shape
    tokens = torch.empty((b, p_ + 1, c), dtype=torch.float32, device=patch.device)
    _ck(L.load().gdl_vit_assemble_tokens(L.ptr(patch), L.dt_code(patch.dtype), L.ptr(pos), L.ptr(cls), L.ptr(tokens), b, p_, c,
                                         L.stream_ptr()))
    return tokens


def vit_extract_feature(tokens, dtype):
    """fp32 tokens (B,P+1,C) -> (B,P,C) in `dtype` without the cls token"""
    b, p1, c = tokens.shape
    feat = torch.empty((b, p1 - 1, c), dtype=dtype, device=tokens.device)
    _ck(L.load().gdl_vit_extract_feature(L.ptr(tokens), L.ptr(feat), L.dt_code(dtype), b, p1 - 1, c, L.stream_ptr()))
    return feat


def vit_feature_grad(dfeat: torch.Tensor, g: torch.Tensor | None) -> torch.Tensor:
    """backward of vit_extract_feature: dfeat (B,P,C) is added to rows 1.. of the fp32 stream gradient g (B,P+1,C);
    g=None starts it (cls row zero)."""
    b, p, c = dfeat.shape
    init = g is None
    if init:
        g = torch.empty((b, p + 1, c), dtype=torch.float32, device=dfeat.device)
    if not dfeat.is_contiguous() or not g.is_contiguous() or tuple(g.shape) != (b, p + 1, c):
        raise ValueError("vit_feature_grad: contiguous (B,P,C) / (B,P+1,C) tensors expected")
    _ck(L.load().gdl_vit_feature_grad(L.ptr(dfeat), L.dt_code(dfeat.dtype), L.ptr(g), b, p, c, int(init), L.stream_ptr()))
    return g


def gelu_fwd(x: torch.Tensor) -> torch.Tensor:
    """exact GELU of a contiguous 16-bit tensor (the pre-activation stays with the caller for gelu_bwd)"""
    if not x.is_contiguous():
        raise ValueError("gelu_fwd: contiguous input expected")
    y = torch.empty_like(x)
    _ck(L.load().gdl_gelu_fwd(L.ptr(x), L.ptr(y), L.dt_code(x.dtype), x.numel(), L.stream_ptr()))
    return y


def gelu_bwd(dy: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    if not dy.is_contiguous() or not pre.is_contiguous() or dy.shape != pre.shape:
        raise ValueError("gelu_bwd: contiguous tensors of one shape expected")
    dpre = torch.empty_like(pre)
    _ck(L.load().gdl_gelu_bwd(L.ptr(dy), L.ptr(pre), L.ptr(dpre), L.dt_code(pre.dtype), pre.numel(), L.stream_ptr()))
    return dpre


def layerscale_add(res: torch.Tensor, u: torch.Tensor, gamma: torch.Tensor, sscale: torch.Tensor | None = None,
                   rows_per_sample: int = 0) -> torch.Tensor:
    """fp32 stream (M,C) + s(row) * gamma * u (16-bit (M,C)) -> new fp32 stream; sscale (B,) = DropPath factors."""
    m, c = u.shape
    if tuple(res.shape) != (m, c) or not res.is_contiguous() or not u.is_contiguous():
        raise ValueError("layerscale_add: contiguous (M,C) tensors expected")
    out = torch.empty_like(res)
    _ck(L.load().gdl_layerscale_add(L.ptr(res), L.ptr(u), L.dt_code(u.dtype), L.ptr(gamma), L.ptr(sscale),
                                        int(rows_per_sample), L.ptr(out), m, c, L.stream_ptr()))
    return out


def layerscale_bwd(g: torch.Tensor, u: torch.Tensor, gamma: torch.Tensor, dgamma: torch.Tensor | None,
                   sscale: torch.Tensor | None = None, rows_per_sample: int = 0) -> torch.Tensor:
    """g fp32 (M,C) stream gradient -> du (16-bit) = s * gamma * g; dgamma (fp32 (C,), pre-zeroed) += sum_r s * g * u."""
    m, c = u.shape
    if tuple(g.shape) != (m, c) or not g.is_contiguous() or not u.is_contiguous():
        raise ValueError("layerscale_bwd: contiguous (M,C) tensors expected")
    du = torch.empty_like(u)
    _ck(L.load().gdl_layerscale_bwd(L.ptr(g), L.ptr(u), L.dt_code(u.dtype), L.ptr(gamma), L.ptr(sscale),
                                        int(rows_per_sample), L.ptr(du), L.ptr(dgamma), m, c, L.stream_ptr()))
    return du


def dropout2d_apply(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x 16-bit NHWC (N,H,W,C), mask fp32 (N,C) = keep / (1 - p): y = x * mask[n][c] (forward and backward of Dropout2d)"""
    n, h, w, c, ld = _nhwc_src(x)
    if tuple(mask.shape) != (n, c) or mask.dtype != torch.float32 or not mask.is_contiguous():
        raise ValueError(f"dropout2d: mask must be a contiguous fp32 ({n}, {c}) tensor")
    y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
    _ck(L.load().gdl_dropout2d_apply(L.ptr(x), ld, L.ptr(mask), L.ptr(y), c, L.dt_code(x.dtype), n, h * w, c,
                                         L.stream_ptr()))
    return y


def channel_pool_fwd(xw: torch.Tensor, scores: torch.Tensor, bands: int, want_attn: bool = True):
    """xw 16-bit (N,H,W,bands*E), scores fp32 (N,H,W,16) -> (out 16-bit (N,H,W,E), attn fp32 (N,H,W,16) | None):
    softmax over the bands of the channel-attention logits and the weighted sum of the per-band maps."""
    n, h, w, ce = xw.shape
    e = ce // bands
    if not xw.is_contiguous() or not scores.is_contiguous() or tuple(scores.shape) != (n, h, w, 16) or \
            scores.dtype != torch.float32 or e * bands != ce:
        raise ValueError("channel_pool_fwd: contiguous xw (N,H,W,bands*E) and fp32 scores (N,H,W,16) expected")
    out = torch.empty((n, h, w, e), dtype=xw.dtype, device=xw.device)
    attn = torch.empty((n, h, w, 16), dtype=torch.float32, device=xw.device) if want_attn else None
    _ck(L.load().gdl_channel_pool_fwd(L.ptr(xw), L.ptr(scores), L.ptr(out), L.ptr(attn), L.dt_code(xw.dtype), n * h * w, bands,
                                          e, L.stream_ptr()))
    return out, attn


def channel_pool_bwd(dout: torch.Tensor, xw: torch.Tensor, attn: torch.Tensor, bands: int):
    """-> (dxw 16-bit like xw, dscores 16-bit (N,H,W,16))"""
    n, h, w, ce = xw.shape
    e = ce // bands
    if not dout.is_contiguous() or not xw.is_contiguous() or not attn.is_contiguous() or tuple(dout.shape) != (n, h, w, e):
        raise ValueError("channel_pool_bwd: contiguous dout (N,H,W,E), xw (N,H,W,bands*E), attn (N,H,W,16) expected")
    dxw = torch.empty_like(xw)
    ds = torch.empty((n, h, w, 16), dtype=xw.dtype, device=xw.device)
    _ck(L.load().gdl_channel_pool_bwd(L.ptr(dout), L.ptr(xw), L.ptr(attn), L.ptr(dxw), L.ptr(ds), L.dt_code(xw.dtype), n * h * w,
                                          bands, e, L.stream_ptr()))
    return dxw, ds


def relu_bwd(dy: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    if not dy.is_contiguous() or not y.is_contiguous() or dy.shape != y.shape:
        raise ValueError("relu_bwd: contiguous tensors of one shape expected")
    dx = torch.empty_like(y)
    _ck(L.load().gdl_relu_bwd(L.ptr(dy), L.ptr(y), L.ptr(dx), L.dt_code(y.dtype), y.numel(), L.stream_ptr()))
    return dx
