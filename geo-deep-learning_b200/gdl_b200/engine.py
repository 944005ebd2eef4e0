"""Static-graph execution engine: forward ops that record hand-written backward closures.

torch.autograd never differentiates through the hot path: every op below launches the sm_100a
kernels of libgdlb200.so for its forward AND registers a closure that launches the matching
backward kernels.  Gradients w.r.t. an activation are not summed eagerly: each consumer
registers a *gradient source* (a channel slice of its dgrad output, or the gradient of the
nearest-x2-upsampled copy) on the activation, and the producer's backward reads them all in
one fused `grad_gather` pass (sum + ReLU mask + BatchNorm-backward reductions).

Layout: NHWC 16-bit activations (bf16 or fp16), fp32 parameters / statistics / gradients.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import torch

from . import ops


class Act:
    """An activation (N,H,W,C) + the gradient sources registered by its consumers."""

    __slots__ = ("t", "gsrcs", "needs_grad", "up")

    def __init__(self, t: torch.Tensor, needs_grad: bool = True) -> None:
        self.t = t
        self.gsrcs: list[tuple[torch.Tensor, int]] = []
        self.needs_grad = needs_grad
        self.up: Act | None = None  # nearest-x2 upsampled copy written by the producer

    @property
    def shape(self):
        return self.t.shape

    def all_gsrcs(self) -> list[tuple[torch.Tensor, int]]:
        s = list(self.gsrcs)
        if self.up is not None:
            s += [(t, 1) for t, _ in self.up.gsrcs]
        return s


@dataclass
class BNParams:
    weight: torch.Tensor
    bias: torch.Tensor
    running_mean: torch.Tensor
    running_var: torch.Tensor
    num_batches_tracked: torch.Tensor | None
    eps: float = 1e-5
    momentum: float = 0.1


@dataclass
class RawConv:
    """Raw (pre-normalisation) conv output + what its backward needs."""
    x: torch.Tensor
    srcs: list[Act]
    weight: torch.nn.Parameter
    stride: int
    pad: int
    col: torch.Tensor | None = None  # im2col matrix (strided / narrow-input convs)
    kpad: int = 0
    cin_store: int = 0  # channels per pixel as stored (>= weight.shape[1] when the input is padded)
    wshape: tuple | None = None  # (Cout, Cin, R, S) view of the parameter (nn.Linear weights are (Cout, Cin))
    bias: torch.Tensor | None = None
    pixel_packed: bool = False  # narrow 3x3 conv run as its pixel-packed (block-Toeplitz) equivalent
    stats: torch.Tensor | None = None  # BatchNorm sums [2C] (pivoted on the running mean) produced by the conv itself
    groups: int = 1             # grouped conv (ResNeXt): run as Cout / 64 dense 64 -> 64 convs over channel slices
    cols: list | None = None    # grouped + strided: the im2col matrix of every 64-channel slice
    custom_backward: Callable | None = None  # the producer of x is not one conv of `weight`: it owns the backward of d(x)


@dataclass
class BNState:
    p: BNParams
    scale: torch.Tensor
    shift: torch.Tensor
    mean: torch.Tensor | None = None
    invstd: torch.Tensor | None = None
    count: int = 0


def refresh_packed_weights(wcache: dict) -> bool:
    """The owner of `wcache` changed parameters in place BEHIND torch's version counters (the fused trainer's flat-buffer
    Adam): bring every cached 16-bit operand up to date in place — one gdl_repack_weights launch for the plain packings,
    then the few derived operands (zero-padded dgrad weights, block-Toeplitz widenings, tiled biases) — instead of dropping
    the cache and re-packing ~150-200 weights one launch at a time at the start of the next step.  Cached tensors keep their
    addresses, so a captured CUDA graph stays valid.  Returns False when the cache holds nothing refreshable."""
    items = [(k, v) for k, v in wcache.items() if isinstance(k, tuple) and isinstance(v, tuple) and len(v) == 3]
    if not items:
        return False
    with torch.no_grad():
        for _, v in items:  # slice-dense masters of grouped convs: rebuilt from the parameter before anything is packed from them
            if v[2][0] == "slice_dense":
                _, w, rows, idx = v[2]
                v[1][rows, idx] = w.detach().to(v[1].dtype)
    plain = [(v[2][1], v[1], *v[2][1].shape, v[2][2], v[2][3]) for _, v in items if v[2][0] == "plain"]
    # plain entries of one cache share the engine's dtype; group defensively
    by_dtype: dict = {}
    for e in plain:
        by_dtype.setdefault(e[1].dtype, []).append(e)
    for dt, entries in by_dtype.items():
        sig = tuple((e[0].data_ptr(), e[1].data_ptr()) for e in entries)
        pk = ("__repack_plan__", dt)
        old = wcache.get(pk)
        plan = old[1] if old is not None and old[0] == sig else None
        plan = ops.repack_weights(entries, plan)
        wcache[pk] = (sig, plan)
    for k, v in items:  # up to date with their masters as they are NOW (slice-dense masters were just rewritten in place)
        if v[2][0] == "plain":
            wcache[k] = (v[2][1]._version, v[1], v[2])
    with torch.no_grad():
        for _, v in items:  # padded dgrad weights first: widenings may read them
            if v[2][0] == "dgrad_pad":
                _, wv, wpad, cout = v[2]
                wpad[:cout].copy_(wv)
                ops.pack_conv_weight(wpad, v[1].dtype, 1, out=v[1])
        for _, v in items:
            if v[2][0] == "wide":
                _, src, co, ci, r, f = v[2]
                ops.widen_conv_weight(src, co, ci, r, f, out=v[1])
            elif v[2][0] == "tiled_bias":
                _, bias, f = v[2]
                v[1].copy_(bias.detach().to(v[1].dtype).repeat(f))
    return True


class Engine:
    def __init__(self, dtype: torch.dtype = torch.bfloat16, training: bool = True, wcache: dict | None = None,
                 grad_dst: dict[int, torch.Tensor] | None = None, sync_bn_group=None,
                 acc_dtype: torch.dtype = torch.float32, sync_bn_exchange=None) -> None:
        """wcache: persistent dict for packed 16-bit weights (owner clears it when it updates parameters
        behind torch's back); grad_dst: id(param) -> pre-zeroed fp32 tensor the gradient is written into
        (e.g. a view of a flat gradient buffer); sync_bn_group: torch.distributed process group for
        SyncBatchNorm statistics (None = per-rank statistics)."""
        self.dtype = dtype
        self.acc_dtype = acc_dtype  # statistics / gradients / logits (fp32; float64 only in CPU host-logic tests)
        self.training = training
        self.tape: list[Callable[[], None]] = []
        self._wcache: dict = wcache if wcache is not None else {}
        self.grad_dst = grad_dst or {}
        self.sync_bn_group = sync_bn_group
        self.sync_bn_exchange = sync_bn_exchange  # ops.P2PExchange: the statistics travel over NVLink peer memory instead of NCCL
        self.param_grads: dict[int, torch.Tensor] = {}
        self.on_progress = None  # optional callback(engine) after every backward closure
        self._head = None

    # ------------------------------------------------------------------ parameters
    def packed(self, w: torch.Tensor, mode: int, ld: int = 0, wshape: tuple | None = None) -> torch.Tensor:
        """16-bit operand of a weight, cached until the parameter is modified in place (torch's version counter) — or
        refreshed in place by refresh_packed_weights() when the owner updates parameters behind that counter."""
        key = (w.data_ptr(), mode, ld, self.dtype)
        ver = w._version
        hit = self._wcache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        wv = w.detach()
        if wshape is not None:
            wv = wv.view(wshape)
        out = ops.pack_conv_weight(wv, self.dtype, mode, ld)
        self._wcache[key] = (ver, out, ("plain", wv, mode, out.stride(0)))
        return out

    def packed_wide(self, w: torch.Tensor, mode: int, f: int, wshape: tuple, coutp: int = 0) -> torch.Tensor:
        """Block-Toeplitz operand of a narrow Rx3 conv run on f-pixel packed pixels (ops.widen_conv_weight);
        mode 1 = the dgrad operand, whose `coutp` input channels may be zero-padded beyond Cout."""
        key = (w.data_ptr(), "wide", mode, f, coutp, self.dtype)
        hit = self._wcache.get(key)
        if hit is not None and hit[0] == w._version:
            return hit[1]
        cout, cin, r, _ = wshape
        if mode == 0:
            src, co, ci = self.packed(w, 0, 0, wshape), cout, cin
        else:
            src, co, ci = self._dgrad_weight(w, coutp or cout, wshape), cin, coutp or cout
        out = ops.widen_conv_weight(src, co, ci, r, f)
        self._wcache[key] = (w._version, out, ("wide", src, co, ci, r, f))
        return out

    def _tiled_bias(self, bias: torch.Tensor, f: int) -> torch.Tensor:
        key = (bias.data_ptr(), "tiled_bias", f)
        hit = self._wcache.get(key)
        if hit is not None and hit[0] == bias._version:
            return hit[1]
        out = bias.detach().to(self.acc_dtype).repeat(f)
        self._wcache[key] = (bias._version, out, ("tiled_bias", bias, f))
        return out

    @staticmethod
    def _pixel_packable(srcs: list[Act], wshape: tuple, stride: int, pad: int) -> bool:
        """Single dense source, 3x3 / stride 1 / pad 1, 16 or 32 channels in and out, width divisible by 4: run the
        conv on 64-channel packed pixels (128-byte rows, N >= 32) instead of 32/64-byte rows with N = 16."""
        cout, cin, r, s = wshape
        if len(srcs) != 1 or stride != 1 or pad != 1 or r != 3 or s != 3:
            return False
        if cin not in (16, 32) or cout > 32:
            return False
        t = srcs[0].t
        return t.is_contiguous() and t.shape[3] == cin and t.shape[2] % 4 == 0 and ops.option("pixel_pack") != 0

    def grad_buffer(self, p: torch.Tensor, zero: bool) -> torch.Tensor:
        """Destination for d(loss)/d(p): the registered (pre-zeroed) view, else a fresh tensor."""
        k = id(p)
        if k in self.param_grads:
            raise NotImplementedError("parameter used twice in one graph")
        dst = self.grad_dst.get(k)
        if dst is None:
            dst = torch.zeros_like(p, memory_format=torch.contiguous_format) if zero else torch.empty_like(
                p, memory_format=torch.contiguous_format)
        self.param_grads[k] = dst
        return dst

    # ------------------------------------------------------------------ convolution
    def conv_raw(self, srcs: list[Act], weight: torch.nn.Parameter, stride: int, pad: int,
                 bias: torch.Tensor | None = None, out_dtype: torch.dtype | None = None,
                 relu: bool = False, wshape: tuple | None = None, residual: torch.Tensor | None = None,
                 stats_for: "BNParams | None" = None) -> RawConv:
        """Convolution / linear layer.  `wshape` views the parameter as (Cout, Cin, R, S) (nn.Linear);
        `residual` (N,H,W,Cout) is added in the GEMM epilogue (fp32 residual stream).
        stats_for: the BatchNorm that follows (training mode): the conv then also produces that layer's batch sums
        (pivoted on its running mean) — from its epilogue where the kernel can — and bn_prepare skips the statistics pass."""
        wshape = tuple(wshape) if wshape is not None else tuple(weight.shape)
        cout, cin, r, s = wshape
        sums = pivot = None
        if stats_for is not None and self.training and ops.option("bn_fused"):
            sums = torch.empty(2 * cout, dtype=self.acc_dtype, device=srcs[0].t.device)
            pivot = stats_for.running_mean
        stored = sum(a.t.shape[3] for a in srcs)
        direct = stride == 1 and all(a.t.shape[3] % 16 == 0 for a in srcs) and stored == cin
        if direct and residual is None and self._pixel_packable(srcs, wshape, stride, pad):
            a = srcs[0]
            n, h, wd, _ = a.t.shape
            f = 64 // cin
            wide_sums = wide_piv = None
            if sums is not None:  # the packed conv has f * Cout pseudo-channels: f copies of every real channel
                wide_sums = torch.empty(2 * f * cout, dtype=self.acc_dtype, device=a.t.device)
                wide_piv = pivot.repeat(f)
            x = ops.conv2d_fwd([a.t.view(n, h, wd // f, f * cin)], self.packed_wide(weight, 0, f, wshape), f * cout,
                               r, s, pad, pad, out_dtype=out_dtype, relu=relu,
                               bias=self._tiled_bias(bias, f) if bias is not None else None, alg_scale=1.0 / f,
                               bn_sums=wide_sums, bn_pivot=wide_piv)
            if sums is not None:
                torch.sum(wide_sums.view(2, f, cout), dim=1, out=sums.view(2, cout))  # fixed order: reproducible
            return RawConv(x.view(n, h, wd, cout), srcs, weight, stride, pad, cin_store=stored, wshape=wshape, bias=bias,
                           pixel_packed=True, stats=sums)
        if direct:
            wp = self.packed(weight, 0, 0, wshape)
            x = ops.conv2d_fwd([a.t for a in srcs], wp, cout, r, s, pad, pad, out_dtype=out_dtype, bias=bias,
                               relu=relu, residual=residual, bn_sums=sums, bn_pivot=pivot)
            return RawConv(x, srcs, weight, stride, pad, cin_store=stored, wshape=wshape, bias=bias, stats=sums)
        if len(srcs) != 1:
            raise NotImplementedError("strided / narrow-input convs take a single source")
        a = srcs[0]
        k = r * s * cin
        kpad = (k + 63) // 64 * 64
        col = ops.im2col(a.t, cin, r, s, stride, pad, kpad)
        wp = self.packed(weight, 0, kpad, wshape)
        x = ops.conv2d_fwd([col], wp, cout, 1, 1, 0, 0, out_dtype=out_dtype, bias=bias, relu=relu, residual=residual,
                           alg_scale=k / kpad, bn_sums=sums, bn_pivot=pivot)
        return RawConv(x, srcs, weight, stride, pad, col=col if self.training else None, kpad=kpad, cin_store=stored,
                       wshape=wshape, bias=bias, stats=sums)

    def conv_backward(self, rc: RawConv, dx: torch.Tensor, dgrad_residual: torch.Tensor | None = None) -> None:
        """dx = gradient w.r.t. the raw conv output (N,Ho,Wo,Cout'), Cout' >= Cout zero padded.
        Weight (and bias) gradients go to grad_buffer; input gradients are registered as gradient sources on
        the source activations.  `dgrad_residual` is added to the (single-source) input gradient in the dgrad
        GEMM epilogue (sum of two gradient paths without an extra pass)."""
        if rc.custom_backward is not None:
            return rc.custom_backward(dx)
        if rc.groups > 1:
            return self._grouped_backward(rc, dx)
        w = rc.weight
        cout, cin, r, s = rc.wshape
        coutp = dx.shape[3]
        dev = dx.device
        if rc.bias is not None and rc.bias.requires_grad:
            sums = torch.empty(2 * coutp, dtype=self.acc_dtype, device=dev)
            ops.bn_stats(dx, sums)  # column sums (no pivot)
            self.grad_buffer(rc.bias, False).copy_(sums[:cout])
        if rc.pixel_packed and coutp in (16, 32) and dgrad_residual is None and dx.is_contiguous():
            a = rc.srcs[0]
            n, h, wd, _ = a.t.shape
            if w.requires_grad:
                f = 64 // max(cin, coutp)
                dw = torch.zeros((f * coutp, r * s * f * cin), dtype=self.acc_dtype, device=dev)
                ops.conv2d_wgrad([a.t.view(n, h, wd // f, f * cin)], dx.view(n, h, wd // f, f * coutp), r, s, rc.pad,
                                 rc.pad, dw, alg_scale=cout / (f * coutp))
                ops.fold_widened_wgrad(dw, self.grad_buffer(w, False).view(cout, cin, r, s), f)
            if a.needs_grad:
                f = 64 // coutp
                dcat = ops.conv2d_fwd([dx.view(n, h, wd // f, f * coutp)], self.packed_wide(w, 1, f, rc.wshape, coutp),
                                      f * cin, r, s, r - 1 - rc.pad, s - 1 - rc.pad, alg_scale=cout / (f * coutp))
                a.gsrcs.append((dcat.view(n, h, wd, cin), 0))
            return
        if rc.col is None:
            if w.requires_grad:
                if r == 1 and s == 1 and coutp == cout:
                    # OIHW == [Cout][(r,s,c)] for a pointwise conv: accumulate straight into the gradient
                    ops.conv2d_wgrad([a.t for a in rc.srcs], dx, 1, 1, 0, 0, self.grad_buffer(w, True).view(cout, cin))
                else:
                    dw = torch.zeros((coutp, r * s * cin), dtype=self.acc_dtype, device=dev)
                    ops.conv2d_wgrad([a.t for a in rc.srcs], dx, r, s, rc.pad, rc.pad, dw, alg_scale=cout / coutp)
                    ops.unpack_conv_wgrad(dw, self.grad_buffer(w, False).view(cout, cin, r, s), r * s * cin)
            if any(a.needs_grad for a in rc.srcs):
                wt = self._dgrad_weight(w, coutp, rc.wshape)
                dcat = ops.conv2d_fwd([dx], wt, cin, r, s, r - 1 - rc.pad, s - 1 - rc.pad, residual=dgrad_residual,
                                      alg_scale=cout / coutp)
                off = 0
                for a in rc.srcs:
                    c = a.t.shape[3]
                    if a.needs_grad:
                        a.gsrcs.append((dcat[..., off:off + c], 0))
                    off += c
        else:
            a = rc.srcs[0]
            if w.requires_grad:
                dw = torch.zeros((coutp, rc.kpad), dtype=self.acc_dtype, device=dev)
                ops.conv2d_wgrad([rc.col], dx, 1, 1, 0, 0, dw, alg_scale=(cout * r * s * cin) / (coutp * rc.kpad))
                ops.unpack_conv_wgrad(dw, self.grad_buffer(w, False).view(cout, cin, r, s), rc.kpad)
            if a.needs_grad:
                if coutp != cout:
                    raise NotImplementedError("padded-output dgrad through im2col")
                wt = self.packed(w, 2, 0, rc.wshape)  # [(r,s,c)][Cout]
                dcol = ops.conv2d_fwd([dx], wt, r * s * cin, 1, 1, 0, 0)  # (stored channels == cin here)
                n, h, wd, c = a.t.shape
                a.gsrcs.append((ops.col2im(dcol, n, h, wd, c, r, s, rc.stride, rc.pad), 0))

    def _dgrad_weight(self, w: torch.Tensor, coutp: int, wshape: tuple) -> torch.Tensor:
        cout = wshape[0]
        if coutp == cout:
            return self.packed(w, 1, 0, wshape)
        key = (w.data_ptr(), "dgrad_pad", coutp, self.dtype)
        hit = self._wcache.get(key)
        if hit is not None and hit[0] == w._version:
            return hit[1]
        wpad = torch.zeros((coutp, *wshape[1:]), dtype=self.acc_dtype, device=w.device)
        wpad[:cout] = w.detach().view(wshape)
        out = ops.pack_conv_weight(wpad, self.dtype, 1)
        self._wcache[key] = (w._version, out, ("dgrad_pad", w.detach().view(wshape), wpad, cout))
        return out

    # ------------------------------------------------------------------ grouped convolution (ResNeXt)
    GROUP_SLICE = 64  # channels per dense slice: 64 / (channels per group) groups side by side, block-diagonal weights

    def _slice_dense(self, w: torch.Tensor, groups: int):
        """Grouped weight (C, C/groups, R, S) -> fp32 "slice-dense" master (C, 64, R, S): output channel o reads the 64 input
        channels of ITS 64-channel slice; entries outside o's group are zero.  Rows [64 j, 64 j + 64) are then the ordinary
        dense OIHW weight of slice j (so the existing packers, kernels and the batched refresh apply unchanged)."""
        key = (w.data_ptr(), "slice_dense", groups)
        hit = self._wcache.get(key)
        if hit is not None and hit[0] == w._version:
            return hit[1], hit[2][2], hit[2][3]
        c, cg, r, s = w.shape
        sl = self.GROUP_SLICE
        if c != cg * groups or c % sl or sl % cg:
            raise NotImplementedError(f"grouped conv: {c} channels in {groups} groups of {cg} (needs C % 64 == 0, 64 % (C / groups) == 0, C_in == C_out)")
        if hit is not None:
            ws, rows, idx = hit[1], hit[2][2], hit[2][3]
        else:
            ws = torch.zeros((c, sl, r, s), dtype=self.acc_dtype, device=w.device)
            o = torch.arange(c, device=w.device)
            rows = o.view(c, 1)
            idx = ((o % sl) // cg * cg).view(c, 1) + torch.arange(cg, device=w.device).view(1, cg)
        with torch.no_grad():
            ws[rows, idx] = w.detach().to(ws.dtype)  # in place: packed operands cached against ws._version are re-packed
        self._wcache[key] = (w._version, ws, ("slice_dense", w, rows, idx))
        return ws, rows, idx

    def conv_raw_grouped(self, a: Act, weight: torch.nn.Parameter, groups: int, stride: int, pad: int) -> RawConv:
        """nn.Conv2d(C, C, 3, stride, padding=1, groups=groups, bias=False) (torchvision ResNeXt Bottleneck.conv2) as C / 64
        dense 64 -> 64 convolutions over channel slices of the NHWC tensors (block-diagonal inside a slice: a factor
        64 / (C / groups) of padded MMA work, none for 32x8d's layer4)."""
        c, cg, r, s = weight.shape
        ws, _, _ = self._slice_dense(weight, groups)
        sl = self.GROUP_SLICE
        n, h, wd, cs = a.t.shape
        if cs != c:
            raise ValueError("grouped conv: input channel count mismatch")
        ho, wo = (h + 2 * pad - r) // stride + 1, (wd + 2 * pad - s) // stride + 1
        x = torch.empty((n, ho, wo, c), dtype=self.dtype, device=a.t.device)
        scale = cg / sl
        cols = None
        if stride == 1:
            wp = self.packed(ws, 0)  # [C][R*S*64]: rows of slice j = its dense operand
            for j in range(c // sl):
                ops.conv2d_fwd([a.t[..., j * sl:(j + 1) * sl]], wp[j * sl:(j + 1) * sl], sl, r, s, pad, pad,
                               out=x[..., j * sl:(j + 1) * sl], alg_scale=scale)
        else:
            k = r * s * sl
            wp = self.packed(ws, 0, k)
            cols = []
            for j in range(c // sl):
                col = ops.im2col(a.t[..., j * sl:(j + 1) * sl], sl, r, s, stride, pad, k)
                ops.conv2d_fwd([col], wp[j * sl:(j + 1) * sl], sl, 1, 1, 0, 0, out=x[..., j * sl:(j + 1) * sl], alg_scale=scale)
                if self.training:
                    cols.append(col)
        return RawConv(x, [a], weight, stride, pad, cin_store=c, wshape=(c, cg, r, s), groups=groups, cols=cols, kpad=r * s * sl)

    def _grouped_backward(self, rc: RawConv, dx: torch.Tensor) -> None:
        w = rc.weight
        c, cg, r, s = rc.wshape
        sl = self.GROUP_SLICE
        a = rc.srcs[0]
        ws, rows, idx = self._slice_dense(w, rc.groups)
        nsl = c // sl
        dev = dx.device
        scale = cg / sl
        n, h, wd, _ = a.t.shape
        if w.requires_grad:
            dws = torch.zeros((c, r * s * sl), dtype=self.acc_dtype, device=dev)
            for j in range(nsl):
                sj = slice(j * sl, (j + 1) * sl)
                if rc.cols is None:
                    ops.conv2d_wgrad([a.t[..., sj]], dx[..., sj], r, s, rc.pad, rc.pad, dws[sj], alg_scale=scale)
                else:
                    ops.conv2d_wgrad([rc.cols[j]], dx[..., sj], 1, 1, 0, 0, dws[sj], alg_scale=scale)
            dense = torch.empty((c, sl, r, s), dtype=self.acc_dtype, device=dev)
            ops.unpack_conv_wgrad(dws, dense, r * s * sl)
            self.grad_buffer(w, False).copy_(dense[rows, idx])  # the diagonal blocks = the grouped weight's gradient
        if a.needs_grad:
            dcat = torch.empty((n, h, wd, c), dtype=self.dtype, device=dev)
            if rc.cols is None:
                for j in range(nsl):
                    sj = slice(j * sl, (j + 1) * sl)
                    wt = self.packed(ws[sj], 1)  # [64][(R-1-r,S-1-s,k)] of slice j
                    ops.conv2d_fwd([dx[..., sj]], wt, sl, r, s, r - 1 - rc.pad, s - 1 - rc.pad, out=dcat[..., sj], alg_scale=scale)
            else:
                for j in range(nsl):
                    sj = slice(j * sl, (j + 1) * sl)
                    wt = self.packed(ws[sj], 2)  # [(r,s,c)][64]
                    dcol = ops.conv2d_fwd([dx[..., sj]], wt, r * s * sl, 1, 1, 0, 0, alg_scale=scale)
                    dslice = ops.col2im(dcol, n, h, wd, sl, r, s, rc.stride, rc.pad)
                    dcat[..., sj].copy_(dslice)
            a.gsrcs.append((dcat, 0))

    # ------------------------------------------------------------------ batch norm
    def _allreduce(self, t: torch.Tensor) -> None:
        if self.sync_bn_exchange is not None:
            self.sync_bn_exchange.all_reduce_(t)
            return
        import torch.distributed as dist
        dist.all_reduce(t, group=self.sync_bn_group)

    def bn_prepare(self, rc: RawConv, p: BNParams) -> BNState:
        c = rc.x.shape[3]
        dev = rc.x.device
        buf = torch.empty((4, c), dtype=self.acc_dtype, device=dev)
        scale, shift, mean, invstd = buf[0], buf[1], buf[2], buf[3]
        if self.training:
            if rc.stats is not None:
                sums = rc.stats  # produced by the conv (its epilogue, or the statistics pass the library ran itself)
            else:
                sums = torch.empty(2 * c, dtype=self.acc_dtype, device=dev)
                # pivot = running mean: the same on every rank, so partial sums add up across ranks
                ops.bn_stats(rc.x, sums, p.running_mean)
            count = ops._rows(rc.x)
            if self.sync_bn_group is not None:
                self._allreduce(sums)
                count *= self.sync_bn_group.size()
            ops.bn_finalize(p.running_mean, sums, count, p.weight, p.bias, p.eps, p.momentum, p.running_mean,
                            p.running_var, scale, shift, mean, invstd)
            if p.num_batches_tracked is not None:
                p.num_batches_tracked += 1
            st = BNState(p, scale, shift, mean, invstd)
            st.count = count
            return st
        ops.bn_eval_coeffs(p.weight, p.bias, p.running_mean, p.running_var, p.eps, scale, shift)
        return BNState(p, scale, shift)

    def bn_act(self, rc: RawConv, bn: BNState, *, relu: bool = True, residual: Act | None = None,
               res_branch: tuple[RawConv, BNState] | None = None, want_up: bool = False,
               want_plain: bool = True) -> Act:
        """y = act(bn(x) [+ residual | + bn_d(x_d)]); optionally also the nearest-x2 upsampled copy."""
        n, h, w, c = rc.x.shape
        dev = rc.x.device
        y = torch.empty((n, h, w, c), dtype=self.dtype, device=dev) if (want_plain or self.training) else None
        yu = torch.empty((n, 2 * h, 2 * w, c), dtype=self.dtype, device=dev) if want_up else None
        res = rscale = rshift = None
        if residual is not None:
            res = residual.t
        elif res_branch is not None:
            res, rscale, rshift = res_branch[0].x, res_branch[1].scale, res_branch[1].shift
        ops.bn_apply(rc.x, bn.scale, bn.shift, res=res, rscale=rscale, rshift=rshift, relu=relu, y=y, y_up=yu)
        out = Act(y)
        if want_up:
            out.up = Act(yu)
        if self.training:
            self.tape.append(lambda: self._bn_act_backward(out, rc, bn, relu, residual, res_branch))
        return out

    def _bn_act_backward(self, out: Act, rc: RawConv, bn: BNState, relu: bool, residual: Act | None,
                         res_branch: tuple[RawConv, BNState] | None) -> None:
        srcs = out.all_gsrcs()
        if not srcs:
            return
        x = rc.x
        c = x.shape[3]
        dev = x.device
        g = torch.empty(x.shape, dtype=self.dtype, device=dev)
        sums = torch.empty(2 * c, dtype=self.acc_dtype, device=dev)
        ops.grad_gather(srcs, x.shape, self.dtype, y=out.t if relu else None, x=x, mean=bn.mean, invstd=bn.invstd,
                        g=g, sums=sums)
        out.gsrcs.clear()
        if out.up is not None:
            out.up.gsrcs.clear()
        dx = torch.empty(x.shape, dtype=self.dtype, device=dev)
        self._bn_param_bwd(g, x, bn, sums, dx)
        self.conv_backward(rc, dx)
        if residual is not None and residual.needs_grad:
            residual.gsrcs.append((g, 0))
        elif res_branch is not None:
            rcd, bnd = res_branch
            sums_d = torch.empty(2 * c, dtype=self.acc_dtype, device=dev)
            ops.grad_gather([(g, 0)], x.shape, self.dtype, x=rcd.x, mean=bnd.mean, invstd=bnd.invstd, sums=sums_d)
            dxd = torch.empty(x.shape, dtype=self.dtype, device=dev)
            self._bn_param_bwd(g, rcd.x, bnd, sums_d, dxd)
            self.conv_backward(rcd, dxd)

    def _bn_param_bwd(self, g, x, bn: BNState, sums, dx) -> None:
        p = bn.p
        dgamma = self.grad_buffer(p.weight, False) if p.weight.requires_grad else None
        dbeta = self.grad_buffer(p.bias, False) if p.bias.requires_grad else None
        if self.sync_bn_group is not None:
            # parameter gradients come from the LOCAL sums (DDP averages them afterwards, as
            # torch SyncBatchNorm does); dx needs the global ones.
            ops.bn_param_grads(sums, dgamma, dbeta)
            self._allreduce(sums)
            dgamma = dbeta = None
        ops.bn_bwd_apply(g, x, bn.mean, bn.invstd, p.weight, sums, dx, dgamma, dbeta, False, bn.count)

    # ------------------------------------------------------------------ pooling
    def maxpool3x3s2(self, a: Act) -> Act:
        y, idx = ops.maxpool3x3s2_fwd(a.t, self.training and a.needs_grad)
        out = Act(y)
        if self.training and a.needs_grad:
            n, h, w, c = a.t.shape

            def bwd() -> None:
                srcs = out.all_gsrcs()
                if not srcs:
                    return
                if len(srcs) == 1 and srcs[0][1] == 0 and srcs[0][0].stride(2) == c:
                    dy = srcs[0][0]
                else:
                    dy = torch.empty(y.shape, dtype=self.dtype, device=y.device)
                    ops.grad_gather(srcs, y.shape, self.dtype, g=dy)
                out.gsrcs.clear()
                a.gsrcs.append((ops.maxpool3x3s2_bwd(dy, idx, h, w), 0))

            self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ generic tape ops
    def collect_grad(self, a: Act) -> torch.Tensor | None:
        """Sum of the gradient sources registered on `a` as one dense 16-bit tensor (None if there are none)."""
        srcs = a.all_gsrcs()
        if not srcs:
            return None
        n, h, w, c = a.t.shape
        if len(srcs) == 1 and srcs[0][1] == 0 and srcs[0][0].stride(2) == c:
            g = srcs[0][0]
        else:
            g = torch.empty((n, h, w, c), dtype=self.dtype, device=a.t.device)
            ops.grad_gather(srcs, a.t.shape, self.dtype, g=g)
        a.gsrcs.clear()
        if a.up is not None:
            a.up.gsrcs.clear()
        return g

    def bilinear(self, a: Act, ho: int, wo: int) -> Act:
        """F.interpolate(mode="bilinear", align_corners=False) to (ho, wo); identity sizes are passed through."""
        n, h, w, c = a.t.shape
        if (h, w) == (ho, wo):
            return a
        out = Act(ops.bilinear_fwd(a.t, ho, wo))
        if self.training and a.needs_grad:
            def bwd() -> None:
                g = self.collect_grad(out)
                if g is not None:
                    a.gsrcs.append((ops.bilinear_bwd(g, h, w), 0))
            self.tape.append(bwd)
        return out

    def add(self, a: Act, b: Act) -> Act:
        out = Act(ops.add_nhwc(a.t, b.t))
        if self.training:
            def bwd() -> None:
                g = self.collect_grad(out)
                if g is not None:
                    if a.needs_grad:
                        a.gsrcs.append((g, 0))
                    if b.needs_grad:
                        b.gsrcs.append((g, 0))
            self.tape.append(bwd)
        return out

    def dropout2d(self, a: Act, p: float, mask: torch.Tensor | None = None) -> Act:
        """nn.Dropout2d(p): identity in eval mode or for p == 0; in training whole (sample, channel) planes are zeroed
        and the rest scaled by 1 / (1 - p).  `mask` (N, C) overrides the draw (tests)."""
        if not self.training or (p <= 0.0 and mask is None):
            return a
        n, _, _, c = a.t.shape
        if mask is None:
            mask = torch.bernoulli(torch.full((n, c), 1.0 - p, dtype=torch.float32, device=a.t.device)) / (1.0 - p)
        mask = mask.to(self.acc_dtype).contiguous()  # fp32 on the GPU
        out = Act(ops.dropout2d_apply(a.t, mask))
        if a.needs_grad:
            def bwd() -> None:
                g = self.collect_grad(out)
                if g is not None:
                    a.gsrcs.append((ops.dropout2d_apply(g, mask), 0))
            self.tape.append(bwd)
        return out

    def adaptive_avgpool(self, a: Act, s: int) -> Act:
        n, h, w, c = a.t.shape
        out = Act(ops.adaptive_avgpool_fwd(a.t, s))
        if self.training and a.needs_grad:
            def bwd() -> None:
                g = self.collect_grad(out)
                if g is not None:
                    a.gsrcs.append((ops.adaptive_avgpool_bwd(g.contiguous(), h, w), 0))
            self.tape.append(bwd)
        return out

    def conv_bn_relu(self, srcs: list[Act], conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d, *, want_up: bool = False) -> Act:
        """ConvModule: conv (stride 1, optional bias) + BatchNorm2d + ReLU"""
        pad = conv.padding[0]
        bnp = BNParams(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.eps,
                       bn.momentum if bn.momentum is not None else 0.1)
        rc = self.conv_raw(srcs, conv.weight, conv.stride[0], pad, bias=conv.bias, stats_for=bnp)
        st = self.bn_prepare(rc, bnp)
        return self.bn_act(rc, st, relu=True, want_up=want_up)

    # ------------------------------------------------------------------ head
    def conv_head(self, a: Act, weight: torch.nn.Parameter, bias: torch.nn.Parameter | None, pad: int) -> torch.Tensor:
        """Conv (+bias) producing fp32 logits (N,H,W,K); its backward takes d(logits) padded to 16 channels."""
        rc = self.conv_raw([a], weight, 1, pad, bias=bias, out_dtype=self.acc_dtype)
        self._head = (rc, bias)
        return rc.x

    def head_backward(self, dlogits16: torch.Tensor) -> None:
        """dlogits16: (N,H,W,16k) 16-bit, channels >= K zero."""
        rc, _ = self._head
        self.conv_backward(rc, dlogits16)  # weight, bias and input gradients

    # ------------------------------------------------------------------ driver
    def backward(self) -> None:
        # pop as we go so each layer's saved activations are released as soon as it is done
        while self.tape:
            self.tape.pop()()
            if self.on_progress is not None:  # the trainer starts the all-reduce of finished gradient buckets here
                self.on_progress(self)
        self._head = None
