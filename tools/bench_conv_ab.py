"""A/B microbenchmark of the conv epilogue variants (and other runtime options) on selected layer shapes."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "geo-deep-learning_b200"))
from gdl_b200 import ops  # noqa: E402

BF = torch.bfloat16
CASES = [
    # name, n, h, w, chans, cout, k, out_dtype, residual
    ("unet 256^2 [64]->448 k3", 32, 256, 256, [64], 448, 3, BF, False),
    ("unet 256^2 [256,64,64,64]->64 k3", 32, 256, 256, [256, 64, 64, 64], 64, 3, BF, False),
    ("unet 128^2 [512,256,256]->256 k3", 32, 128, 128, [512, 256, 256], 256, 3, BF, False),
    ("unet 512^2 [16]->16 k3", 32, 512, 512, [16], 16, 3, BF, False),
    ("unet 128^2 [256]->64 k1", 32, 128, 128, [256], 64, 1, BF, False),
    ("unet 128^2 [64]->256 k1", 32, 128, 128, [64], 256, 1, BF, False),
    ("segf 128^2 [64]->64 k1 f32+res", 16, 128, 128, [64], 64, 1, torch.float32, True),
    ("segf 128^2 [64]->256 k1", 16, 128, 128, [64], 256, 1, BF, False),
    ("segf 128^2 [256]->64 k1 f32+res", 16, 128, 128, [256], 64, 1, torch.float32, True),
    ("segf 128^2 [768]->3072 k1", 16, 128, 128, [768], 3072, 1, BF, False),
    ("segf 64^2 [512]->128 k1 f32+res", 16, 64, 64, [512], 128, 1, torch.float32, True),
]


def time_it(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = []
g = torch.Generator().manual_seed(0)
for name, n, h, w, chans, cout, k, odt, use_res in CASES:
    srcs = [(torch.randn(n, h, w, c, generator=g) * 0.5).to(BF).cuda() for c in chans]
    ctot = sum(chans)
    wt = (torch.randn(cout, ctot, k, k, generator=g) / (ctot * k * k) ** 0.5).cuda()
    wp = ops.pack_conv_weight(wt, BF)
    bias = torch.randn(cout, generator=g).cuda()
    resid = torch.randn(n, h, w, cout, device="cuda") if use_res else None
    out = torch.empty(n, h, w, cout, dtype=odt, device="cuda")
    row = {"case": name, "gflop": 2.0 * n * h * w * cout * k * k * ctot / 1e9}
    outs = {}
    for mode in (0, 1):
        ops.set_option("conv_epilogue", mode)
        fn = lambda: ops.conv2d_fwd(srcs, wp, cout, k, k, k // 2, k // 2, out=out, bias=bias, residual=resid)  # noqa: E731
        ms = time_it(fn)
        row[f"epi{mode}_ms"] = round(ms, 4)
        row[f"epi{mode}_tflops"] = round(row["gflop"] / ms, 1)
        outs[mode] = out.clone()
    row["max_diff"] = (outs[0].float() - outs[1].float()).abs().max().item()
    res.append(row)
    print(row, flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "conv_epilogue_ab.json").write_text(json.dumps(res, indent=1))
