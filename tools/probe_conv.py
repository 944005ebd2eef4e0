"""GPU probe for the tcgen05 implicit-GEMM conv kernels (development tool, run under gpurun).

Runs groups of configurations in child processes (a trapped kernel poisons its CUDA context),
compares with torch fp32 convolution on the same 16-bit-rounded inputs and writes
gpurun_out/probe_conv.json with error statistics and mismatch patterns.
"""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "geo-deep-learning_b200"))

# (name, N, H, W, [src channels], Cout, R, pad, dtype, out_dtype)
FWD = {
    "fwd_gemm": [
        ("g64x16", 1, 1, 128, [64], 16, 1, 0),
        ("g64x64", 1, 1, 128, [64], 64, 1, 0),
        ("g128x128", 1, 1, 256, [128], 128, 1, 0),
        ("g256x256_m1000", 1, 1, 1000, [256], 256, 1, 0),
        ("g3072x768", 2, 32, 32, [3072], 768, 1, 0),
        ("g320x320", 2, 16, 16, [320], 320, 1, 0),
        ("g32x16", 1, 16, 16, [32], 16, 1, 0),
        ("g16x16", 1, 16, 16, [16], 16, 1, 0),
        ("g16x5", 1, 16, 16, [16], 5, 1, 0),
        ("g64x40", 1, 16, 16, [64], 40, 1, 0),
    ],
    "fwd_conv": [
        ("c3_64x64_16", 1, 16, 16, [64], 64, 3, 1),
        ("c3_64x64_32", 2, 32, 32, [64], 64, 3, 1),
        ("c3_128x256_64", 2, 64, 64, [128], 256, 3, 1),
        ("c3_256x128_128w", 1, 8, 256, [256], 128, 3, 1),
        ("c3_multi", 2, 32, 32, [64, 128, 64], 128, 3, 1),
        ("c3_32x16", 1, 64, 64, [32], 16, 3, 1),
        ("c3_16x16", 1, 64, 64, [16], 16, 3, 1),
        ("c3_16x5", 1, 32, 32, [16], 5, 3, 1),
        ("c3_36", 2, 36, 36, [64], 64, 3, 1),
        ("c3_18", 2, 18, 18, [128], 64, 3, 1),
        ("c7_64", 1, 32, 32, [64], 32, 7, 3),
        ("c3_valid", 1, 20, 20, [64], 32, 3, 0),
        ("c3_sliced", 1, 16, 16, [-64], 64, 3, 1),
    ],
}
WGRAD = {
    "wg_gemm": [
        ("wg64x64", 1, 1, 256, [64], 64, 1, 0),
        ("wg128x128", 1, 1, 1024, [128], 128, 1, 0),
        ("wg256x320", 1, 1, 4096, [320], 256, 1, 0),
        ("wg16x16", 1, 1, 512, [16], 16, 1, 0),
        ("wg32x16", 1, 1, 512, [32], 16, 1, 0),
        ("wg768x192", 2, 32, 32, [768], 192, 1, 0),
    ],
    "wg_conv": [
        ("wc3_64x64", 1, 16, 16, [64], 64, 3, 1),
        ("wc3_128x256", 2, 32, 32, [128], 256, 3, 1),
        ("wc3_multi", 2, 32, 32, [64, 128, 64], 128, 3, 1),
        ("wc3_16x16", 1, 64, 64, [16], 16, 3, 1),
        ("wc3_32x16", 1, 64, 64, [32], 16, 3, 1),
        ("wc3_36", 2, 36, 36, [64], 64, 3, 1),
        ("wc7_64", 1, 32, 32, [64], 32, 7, 3),
    ],
}


def run_group(kind, group):
    import torch
    import torch.nn.functional as F
    from gdl_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    results = []
    cfgs = (FWD if kind == "fwd" else WGRAD)[group]
    for (name, n, h, w, chans, cout, r, pad) in cfgs:
        for dt in (torch.bfloat16,):
            rec = {"name": name, "kind": kind, "dtype": str(dt)}
            try:
                g = torch.Generator(device="cpu").manual_seed(1234)
                srcs = []
                for c in chans:
                    if c < 0:  # channel slice of a wider buffer
                        c = -c
                        big = (torch.randn(n, h, w, c + 64, generator=g) * 0.5).to(dt).to(dev)
                        srcs.append(big[..., 32:32 + c])
                    else:
                        srcs.append((torch.randn(n, h, w, c, generator=g) * 0.5).to(dt).to(dev))
                ctot = sum(t.shape[3] for t in srcs)
                x_nchw = torch.cat([t.float() for t in srcs], dim=3).permute(0, 3, 1, 2).contiguous()
                if kind == "fwd":
                    wt = (torch.randn(cout, ctot, r, r, generator=g) / (ctot * r * r) ** 0.5).to(dev)
                    bias = torch.randn(cout, generator=g).to(dev)
                    wp = ops.pack_conv_weight(wt, dt)
                    # pack check
                    ref_pack = wt.permute(0, 2, 3, 1).contiguous().to(dt)
                    rec["pack_ok"] = bool(torch.equal(wp, ref_pack))
                    out = ops.conv2d_fwd(srcs, wp, cout, r, r, pad, pad, out_dtype=torch.float32, bias=bias, relu=True)
                    torch.cuda.synchronize()
                    ref = F.relu(F.conv2d(x_nchw, wp.float().permute(0, 3, 1, 2).contiguous(), bias, padding=pad))
                    ref = ref.permute(0, 2, 3, 1).contiguous()
                    got = out
                    # also bf16 output path
                    out16 = ops.conv2d_fwd(srcs, wp, cout, r, r, pad, pad, relu=False)
                    torch.cuda.synchronize()
                    ref16 = F.conv2d(x_nchw, wp.float().permute(0, 3, 1, 2).contiguous(), None, padding=pad)
                    ref16 = ref16.permute(0, 2, 3, 1).contiguous()
                    e16 = (out16.float() - ref16).abs().max().item()
                    rec["bf16_out_maxerr"] = e16
                    rec["bf16_out_refmax"] = ref16.abs().max().item()
                else:
                    ho, wo = h + 2 * pad - r + 1, w + 2 * pad - r + 1
                    dy = (torch.randn(n, ho, wo, cout, generator=g) * 0.5).to(dt).to(dev)
                    dw = torch.zeros(cout, r, r, ctot, device=dev)
                    ops.conv2d_wgrad(srcs, dy, r, r, pad, pad, dw)
                    torch.cuda.synchronize()
                    # reference: conv of x (as batch=channels) with dy
                    xr = x_nchw.double()
                    dyr = dy.double().permute(0, 3, 1, 2).contiguous()
                    wref = torch.nn.grad.conv2d_weight(xr, (cout, ctot, r, r), dyr, padding=pad)
                    ref = wref.permute(0, 2, 3, 1).contiguous().float()
                    got = dw
                err = (got - ref).abs()
                scale = ref.abs().max().item() + 1e-12
                rec["maxerr"] = err.max().item()
                rec["refmax"] = scale
                rec["relerr"] = rec["maxerr"] / scale
                bad = err > (2e-3 * scale + 1e-4)
                rec["bad_frac"] = bad.float().mean().item()
                rec["ok"] = bool(rec["relerr"] < 2e-3)
                if not rec["ok"]:
                    idx = bad.nonzero()[:12].tolist()
                    rec["bad_idx"] = idx
                    rec["bad_got"] = [got[tuple(i)].item() for i in idx[:6]]
                    rec["bad_ref"] = [ref[tuple(i)].item() for i in idx[:6]]
                    # per-dim pattern: which indices along each dim are bad
                    pat = {}
                    for dim in range(bad.dim()):
                        other = [d for d in range(bad.dim()) if d != dim]
                        v = bad.float().mean(dim=other)
                        pat[f"dim{dim}"] = [round(x, 3) for x in v.tolist()[:64]]
                    rec["pattern"] = pat
            except Exception as e:  # noqa: BLE001
                rec["ok"] = False
                rec["exception"] = f"{type(e).__name__}: {e}"
                results.append(rec)
                print(json.dumps(rec), flush=True)
                break
            results.append(rec)
            print(json.dumps(rec), flush=True)
    return results


def main():
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    if len(sys.argv) >= 3:
        res = run_group(sys.argv[1], sys.argv[2])
        (out_dir / f"probe_{sys.argv[1]}_{sys.argv[2]}.json").write_text(json.dumps(res, indent=1))
        return
    summary = {}
    for kind, groups in (("fwd", FWD), ("wgrad", WGRAD)):
        for gname in groups:
            t0 = time.time()
            try:
                pr = subprocess.run([sys.executable, __file__, kind, gname], capture_output=True, text=True,
                                    timeout=300)
                tail = (pr.stdout[-3000:], pr.stderr[-3000:])
                rc = pr.returncode
            except subprocess.TimeoutExpired as e:
                tail = (str(e.stdout)[-2000:], str(e.stderr)[-2000:])
                rc = -999
            summary[f"{kind}:{gname}"] = {"rc": rc, "sec": round(time.time() - t0, 1), "stdout": tail[0], "stderr": tail[1]}
            print(kind, gname, "rc", rc, flush=True)
    (out_dir / "probe_conv_summary.json").write_text(json.dumps(summary, indent=1))
    # compact table
    for f in sorted(out_dir.glob("probe_*_*.json")):
        for rec in json.loads(f.read_text()):
            print(rec.get("name"), rec.get("kind"), "OK" if rec.get("ok") else "FAIL", rec.get("relerr"),
                  rec.get("bad_frac"), rec.get("exception", ""))


if __name__ == "__main__":
    main()
