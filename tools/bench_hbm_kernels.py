#!/usr/bin/env python
"""Microbenchmark of the HBM-bound kernels added for the widening rows (augmentation + normalise, argmax + confusion,
Dropout2d, GELU, LayerScale, feature-tap gradient): achieved GB/s on ALGORITHMIC bytes (unique input + output bytes of the
launch) against the measured copy bandwidth in MEASURED_PEAKS.json.  CUDA events on the launching stream, 3 warm-up
launches, an L2 flush (256 MB write) between timed launches, working sets larger than the 126 MB L2.

  python tools/bench_hbm_kernels.py [--out profiles/rNN_hbm_kernels.json]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "geo-deep-learning_b200"))


def main() -> None:
    import torch

    from gdl_b200 import _lib, ops
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    _lib.load()
    dev = torch.device("cuda")
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks_path.read_text())["hbm_gbs"] if peaks_path.exists() else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)

    def timed(fn) -> float:
        for _ in range(3):
            fn()
        ms = []
        for _ in range(args.iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        return ms[len(ms) // 2]

    rows = []

    def report(name: str, nbytes: float, fn) -> None:
        ms = timed(fn)
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / peak, 3)})
        print(json.dumps(rows[-1]), flush=True)

    n, t, c, k = 32, 512, 4, 5
    raw = torch.randint(0, 256, (n, t, t, c), dtype=torch.uint8, device=dev, generator=g)
    mask = torch.randint(0, k, (n, t, t), dtype=torch.uint8, device=dev, generator=g)
    mean = torch.full((c,), 0.5, device=dev)
    std = torch.full((c,), 0.2, device=dev)
    pix = n * t * t
    report("normalize_to_nhwc (u8 NHWC -> bf16 NHWC8)", pix * (c + 16), lambda: ops.normalize_to_nhwc(raw, False, torch.bfloat16, 8, mean, std, 255.0))
    for label, op in (("identity", 0), ("hflip", 1), ("rot90", 3), ("crop", 4)):
        p = torch.zeros((n, 6), dtype=torch.int32, device=dev)
        p[:, 0] = op
        p[:, 1] = 1
        p[:, 2:] = torch.tensor([37, 51, 400, 380], dtype=torch.int32, device=dev)
        report(f"augment_normalize [{label}] (u8 + mask -> bf16 NHWC8 + mask)", pix * (c + 1 + 16 + 1),
               lambda p=p: ops.augment_normalize(raw, False, mask, p, torch.bfloat16, 8, mean, std, 255.0))
    logits = torch.randn(n, t, t, k, device=dev, generator=g)
    tgt = mask.long()
    report("argmax_classes (fp32 logits -> int64)", pix * (4 * k + 8), lambda: ops.argmax_classes(logits))
    report("argmax_confusion (fp32 logits + int64 target -> int64 classes + counts)", pix * (4 * k + 8 + 8),
           lambda: ops.argmax_confusion(logits, tgt))
    report("argmax_confusion, counts only, uint8 target", pix * (4 * k + 1), lambda: ops.argmax_confusion(logits, mask, want_classes=False))
    x = torch.randn(16, 128, 128, 768, device=dev, generator=g).bfloat16()
    m = (torch.rand(16, 768, device=dev, generator=g) < 0.9).float() / 0.9
    report("dropout2d_apply (bf16 NHWC)", x.numel() * 4, lambda: ops.dropout2d_apply(x, m))
    b, ntok, d = 16, 1297, 768
    rows_ = b * ntok
    h = torch.randn(rows_, 4 * d, device=dev, generator=g).bfloat16()
    dh = torch.randn(rows_, 4 * d, device=dev, generator=g).bfloat16()
    report("gelu_fwd (bf16)", h.numel() * 4, lambda: ops.gelu_fwd(h))
    report("gelu_bwd (bf16)", h.numel() * 6, lambda: ops.gelu_bwd(dh, h))
    res = torch.randn(rows_, d, device=dev, generator=g)
    u = torch.randn(rows_, d, device=dev, generator=g).bfloat16()
    gam = torch.rand(d, device=dev, generator=g)
    s = torch.ones(b, device=dev)
    dg = torch.zeros(d, device=dev)
    report("layerscale_add (fp32 stream + bf16 branch)", rows_ * d * (4 + 2 + 4), lambda: ops.layerscale_add(res, u, gam, s, ntok))
    report("layerscale_bwd (fp32 g, bf16 u -> bf16 du, dgamma)", rows_ * d * (4 + 2 + 2), lambda: ops.layerscale_bwd(res, u, gam, dg, s, ntok))
    df = torch.randn(b, ntok - 1, d, device=dev, generator=g).bfloat16()
    gs = torch.randn(b, ntok, d, device=dev, generator=g)
    report("vit_feature_grad (accumulate)", b * ntok * d * (2 + 4 + 4), lambda: ops.vit_feature_grad(df, gs))
    bands, e = 4, 64
    xw = torch.randn(16, 128, 128, bands * e, device=dev, generator=g).bfloat16()
    sc = torch.randn(16, 128, 128, 16, device=dev, generator=g)
    pooled, attn = ops.channel_pool_fwd(xw, sc, bands)
    px = 16 * 128 * 128
    report("channel_pool_fwd (4 bands x 64)", px * (bands * e * 2 + 64 + e * 2 + 64), lambda: ops.channel_pool_fwd(xw, sc, bands))
    report("channel_pool_bwd (4 bands x 64)", px * (e * 2 + bands * e * 2 + 64 + bands * e * 2 + 32),
           lambda: ops.channel_pool_bwd(pooled, xw, attn, bands))
    report("relu_bwd (bf16)", xw.numel() * 6, lambda: ops.relu_bwd(xw, xw))
    # ---- the HBM-bound kernels of the UNet++ step at the bench batch: 32 x 256 x 256 x 64 and 32 x 128 x 128 x 256
    del x, h, dh, xw, sc, pooled, attn, logits
    torch.cuda.empty_cache()
    for (nn_, hh, ww, cc) in ((32, 256, 256, 64), (32, 128, 128, 256)):
        xa = torch.randn(nn_, hh, ww, cc, device=dev, generator=g).bfloat16()
        ya = torch.randn(nn_, hh, ww, cc, device=dev, generator=g).bfloat16()
        el = xa.numel()
        piv = torch.zeros(cc, device=dev)
        sums = torch.empty(2 * cc, device=dev)
        tag = f"{nn_}x{hh}x{ww}x{cc}"
        report(f"bn_stats {tag} (bf16 in, 2C sums, ordered)", el * 2, lambda: ops.bn_stats(xa, sums, piv))
        scale, shift = torch.rand(cc, device=dev) + 0.5, torch.randn(cc, device=dev)
        yo = torch.empty_like(xa)
        report(f"bn_apply + ReLU {tag}", el * 4, lambda: ops.bn_apply(xa, scale, shift, relu=True, y=yo))
        yu = torch.empty(nn_, 2 * hh, 2 * ww, cc, dtype=torch.bfloat16, device=dev)
        report(f"bn_apply + ReLU + nearest x2 copy {tag}", el * (2 + 2 + 8), lambda: ops.bn_apply(xa, scale, shift, relu=True, y=yo, y_up=yu))
        del yu
        mean_, inv_ = torch.randn(cc, device=dev) * 0.1, torch.rand(cc, device=dev) + 0.5
        go = torch.empty_like(xa)
        for ns in (1, 2, 4):
            srcs = [(torch.randn(nn_, hh, ww, cc, device=dev, generator=g).bfloat16(), 0) for _ in range(ns)]
            report(f"grad_gather {ns} source(s) + ReLU mask + BN-backward sums {tag}", el * 2 * (ns + 3),
                   lambda srcs=srcs: ops.grad_gather(srcs, xa.shape, torch.bfloat16, y=ya, x=xa, mean=mean_, invstd=inv_, g=go, sums=sums))
            del srcs
        dxo = torch.empty_like(xa)
        gam_ = torch.rand(cc, device=dev)
        report(f"bn_bwd_apply {tag}", el * 6, lambda: ops.bn_bwd_apply(go, xa, mean_, inv_, gam_, sums, dxo, None, None, False))
        del xa, ya, yo, go, dxo
        torch.cuda.empty_cache()
    if args.out:
        import subprocess
        try:
            clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
        except Exception:  # noqa: BLE001
            clk = None
        Path(args.out).write_text(json.dumps({"hbm_peak_GBps": peak, "clocks_after_run": clk, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
