#!/usr/bin/env python
"""gdl_bilinear_sum_fwd at the SegFormer-B2 bench shape (16 tiles of 512²: base 16 x 128 x 128 x 768 + three sources at 64²,
32², 16²) against the launches it replaces (three bilinear_fwd to 128² of the same maps): achieved GB/s on ALGORITHMIC bytes
(unique inputs + the output) vs the measured copy bandwidth.  CUDA events, 3 warm-up launches, a 256 MB L2 flush between
timed launches.  `--once`: three launches only (the command ncu wraps)."""
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
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    _lib.load()
    dev = torch.device("cuda")
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks_path.read_text())["hbm_gbs"] if peaks_path.exists() else 6650.0
    g = torch.Generator(device=dev).manual_seed(0)
    n, h, c = 16, 128, 768
    base = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
    lows = [torch.randn(n, s, s, c, device=dev, generator=g).to(torch.bfloat16) for s in (16, 32, 64)]
    out = torch.empty_like(base)
    if args.once:
        for _ in range(3):
            ops.bilinear_sum_fwd(base, lows, out=out)
        torch.cuda.synchronize()
        return
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn) -> float:
        for _ in range(3):
            fn()
        ms = []
        for _ in range(10):
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
    low_bytes = sum(t.numel() for t in lows) * 2
    ms = timed(lambda: ops.bilinear_sum_fwd(base, lows, out=out))
    nbytes = base.numel() * 2 * 2 + low_bytes
    rows.append({"kernel": "bilinear_sum_vec8_kernel (base + 3 resized sources -> out)", "ms": round(ms, 4),
                 "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1),
                 "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 3)})
    ups = [torch.empty_like(base) for _ in lows]
    ms3 = timed(lambda: [ops.bilinear_fwd(t, h, h, out=u) for t, u in zip(lows, ups)])
    nbytes3 = 3 * base.numel() * 2 + low_bytes
    rows.append({"kernel": "3 x bilinear_fwd_vec8_kernel (the resized maps the reference's op order materialises; read again by the fuse GEMM)",
                 "ms": round(ms3, 4), "algorithmic_MB": round(nbytes3 / 1e6, 1), "GBps": round(nbytes3 / ms3 / 1e6, 1),
                 "frac_of_hbm_peak": round(nbytes3 / ms3 / 1e6 / peak, 3)})
    for r in rows:
        print(json.dumps(r), flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps({"hbm_peak_GBps": peak, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
