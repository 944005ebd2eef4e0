#!/usr/bin/env python
"""Static per-launch counters of the tensor-core kernels, taken on the CPU functional model (tests/hostemu): bytes the TMA engine
moves between L2 and shared memory, tcgen05.mma instructions and MACs, tcgen05.ld count, fp32 global reductions.  These are the
ALGORITHMIC quantities of a launch (what DESIGN.md §4 argues with: L2->SM bytes per tile, wasted MMA work of the narrow layers) —
not timings.  Writes profiles/r01_hostemu_static_counters.json.

  python tools/hostemu_counters.py
"""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "geo-deep-learning_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))

import pytest  # noqa: E402
import torch  # noqa: E402

import hostemu  # noqa: E402

NAMES = ["tma_load_bytes", "tma_store_bytes", "tma_loads", "mma_issued", "mma_macs", "tmem_ld_warp_instr", "mbar_polls_failed", "red_add_f32"]


def counted(fn):
    lib = hostemu.load_lib()
    lib.hostemu_counters_reset()
    fn()
    out = (C.c_ulonglong * 8)()
    lib.hostemu_counters_get(out)
    return dict(zip(NAMES, [int(v) for v in out]))


def main() -> None:
    mp = pytest.MonkeyPatch()
    hostemu.install(mp, torch_convs=False)
    from gdl_b200 import ops
    g = torch.Generator().manual_seed(0)
    bf = torch.bfloat16
    rows = []

    def conv_case(label, n, h, w, cin, cout, r, options=()):
        x = torch.randn(n, h, w, cin, generator=g).to(bf)
        wt = ops.pack_conv_weight(torch.randn(cout, cin, r, r, generator=g) / (cin * r * r) ** 0.5, bf)
        for k, v in options:
            ops.set_option(k, v)
        try:
            c = counted(lambda: ops.conv2d_fwd([x], wt, cout, r, r, r // 2, r // 2))
        finally:
            for k, _ in options:
                ops.set_option(k, 1)
        useful = n * h * w * cout * r * r * cin
        tiles = n * h * ((w + 127) // 128) * ((cout + 255) // 256)
        rows.append({"kernel": label, "shape": f"N{n} {h}x{w} {cin}->{cout} k{r}", **c, "useful_macs": useful,
                     "mma_efficiency": round(useful / max(c["mma_macs"], 1), 3),
                     "tma_load_bytes_per_128px_tile": round(c["tma_load_bytes"] / tiles)})

    # the 64-channel 3x3 layer class of the UNet++ decoder: per-tap tiles vs halo tiles vs the row-streaming kernel
    conv_case("conv_fwd_kernel (per-tap tiles)", 1, 8, 256, 64, 64, 3, (("conv_rows", 0), ("conv_halo", 0)))
    conv_case("conv_fwd_kernel (halo tiles)", 1, 8, 256, 64, 64, 3, (("conv_rows", 0),))
    conv_case("conv3x3_rows_kernel", 1, 8, 256, 64, 64, 3)
    conv_case("conv_fwd_kernel 256->256 (halo tiles)", 1, 4, 256, 256, 256, 3)
    conv_case("conv_fwd_kernel 1x1 256->64", 1, 4, 256, 256, 64, 1)

    def wgrad_case(label, n, h, w, cin, cout, options=()):
        x = torch.randn(n, h, w, cin, generator=g).to(bf)
        dy = torch.randn(n, h, w, cout, generator=g).to(bf)
        dw = torch.zeros(cout, 9 * cin)
        for k, v in options:
            ops.set_option(k, v)
        try:
            c = counted(lambda: ops.conv2d_wgrad([x], dy, 3, 3, 1, 1, dw))
        finally:
            for k, _ in options:
                ops.set_option(k, 1)
        useful = n * h * w * cout * 9 * cin
        rows.append({"kernel": label, "shape": f"N{n} {h}x{w} {cin}->{cout} k3", **c, "useful_macs": useful,
                     "mma_efficiency": round(useful / max(c["mma_macs"], 1), 3)})

    wgrad_case("conv_wgrad_kernel 64->64", 1, 8, 256, 64, 64, (("wgrad_rows", 0),))
    wgrad_case("wgrad3x3_rows_kernel 64->64", 1, 8, 256, 64, 64)
    wgrad_case("conv_wgrad_kernel 128->256", 1, 8, 256, 128, 256)

    # attention: three launches vs the fused kernel (one image, one head, 1024 queries, 256 keys)
    b, n, heads, nk = 1, 1024, 1, 256
    c_ = 64 * heads
    q = torch.randn(b, n, c_, generator=g).to(bf)
    kv2 = torch.randn(b * nk, 2 * c_, generator=g).to(bf)

    def three():
        scores = torch.empty((b, 1, n, heads * nk), dtype=bf)
        ops.conv2d_fwd([q.view(b, 1, n, c_)[..., 0:64]], kv2[:, 0:64], nk, 1, 1, 0, 0, out=scores[..., 0:nk], w_rows_per_img=nk,
                       groups=(heads, 64, 64, nk))
        p = ops.softmax_fwd(scores.view(b, n, heads, nk), 0.125, nk)
        o = torch.empty((b, 1, n, c_), dtype=bf)
        ops.conv2d_fwd([p.view(b, 1, n, heads * nk)[..., 0:nk]], kv2[:, c_:c_ + 64], 64, 1, 1, 0, 0, out=o[..., 0:64], w_rows_per_img=nk,
                       w_mn_major=True, groups=(heads, nk, 64, 64))

    useful = 2 * b * heads * n * nk * 64
    score_bytes = b * heads * n * nk * 2
    for label, fn, hbm in (("attention fwd: q.k^T GEMM + softmax + P.V GEMM", three, 4 * score_bytes),
                           ("sra_attention_fwd_kernel (training: P saved)", lambda: ops.sra_attention_fwd(q, kv2, heads, nk, 0.125, True), score_bytes),
                           ("sra_attention_fwd_kernel (inference)", lambda: ops.sra_attention_fwd(q, kv2, heads, nk, 0.125, False), 0)):
        cc = counted(fn)
        rows.append({"kernel": label, "shape": f"B{b} heads{heads} N{n} keys{nk} d64", **cc, "useful_macs": useful,
                     "mma_efficiency": round(useful / max(cc["mma_macs"], 1), 3), "score_tensor_hbm_bytes": hbm})

    # DOFA-like self-attention (keys streamed in blocks of 128): three launches vs the flash kernel
    n2, heads2 = 1297, 1
    qkv = torch.randn(n2, 3 * 64 * heads2, generator=g).to(bf)

    def three_mha():
        from gdl_b200.models.dofa import _mha
        ops.set_option("mha_flash", 0)
        _mha(qkv, 1, n2, heads2, 64 * heads2, bf)

    lp2 = (n2 + 63) // 64 * 64
    for label, fn, hbm in (("ViT self-attention fwd: q.k^T GEMM + softmax + P.V GEMM", three_mha, 4 * n2 * lp2 * 2),
                           ("sra_attention_kernel, streamed keys (gdl_mha_flash_fwd)", lambda: ops.mha_flash_fwd(qkv, 1, n2, heads2, 0.125), 0)):
        cc = counted(fn)
        useful2 = 2 * heads2 * n2 * n2 * 64
        rows.append({"kernel": label, "shape": f"B1 heads{heads2} N{n2} d64", **cc, "useful_macs": useful2,
                     "mma_efficiency": round(useful2 / max(cc["mma_macs"], 1), 3), "score_tensor_hbm_bytes": hbm})

    out = {"note": "static counters from the CPU functional model of TMA / tcgen05 (tests/hostemu): algorithmic work per launch, not timings; "
                   "tma_load_bytes = L2 -> shared-memory traffic the kernel requests, mma_efficiency = useful MACs / MACs issued",
           "rows": rows}
    path = ROOT / "profiles" / "r01_hostemu_static_counters.json"
    path.write_text(json.dumps(out, indent=1))
    for r_ in rows:
        print(f"{r_['kernel']:52s} {r_['shape']:28s} tma_ld {r_['tma_load_bytes'] / 1e3:9.1f} kB  st {r_['tma_store_bytes'] / 1e3:8.1f} kB  "
              f"mma {r_['mma_issued']:6d}  eff {r_['mma_efficiency']:.2f}  red {r_['red_add_f32']}")
    mp.undo()


if __name__ == "__main__":
    main()
