"""GPU probe: shifted UMMA descriptor start inside a swizzle atom (tools/probe/debug_probe.cu, built here into its own
library together with the product's runtime.cu — the probe is not part of libgdlb200.so)."""
import ctypes
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "geo-deep-learning_b200"))
from gdl_b200 import _lib as L  # noqa: E402

L.load()
_so = ROOT / "tools" / "probe" / "libgdlprobe.so"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                "-I", str(ROOT / "include"), str(ROOT / "tools" / "probe" / "debug_probe.cu"),
                str(ROOT / "geo-deep-learning_b200" / "csrc" / "runtime.cu"), "-o", str(_so)], check=True)
lib = ctypes.CDLL(str(_so))
lib.gdl_debug_shift_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
lib.gdl_debug_shift_probe.restype = ctypes.c_int
g = torch.Generator().manual_seed(0)
res = []
for mode in (0, 1):
    if mode == 0:
        a = torch.randn(144, 64, generator=g).bfloat16().cuda()
        b = torch.randn(64, 64, generator=g).bfloat16().cuda()
    else:
        a = torch.randn(80, 128, generator=g).bfloat16().cuda()
        b = torch.randn(80, 64, generator=g).bfloat16().cuda()
    for bo in (0, 1):
        for shift in range(0, 17):
            out = torch.zeros(128, 64, device="cuda")
            L.check(lib.gdl_debug_shift_probe(L.ptr(a), L.ptr(b), L.ptr(out), mode, shift, bo, L.stream_ptr()))
            torch.cuda.synchronize()
            if mode == 0:
                ref = a[shift:shift + 128].float() @ b.float().t()
            else:
                ref = a[shift:shift + 64].float().t() @ b[:64].float()
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            res.append({"mode": mode, "bo": bo, "shift": shift, "relerr": err})
            print(res[-1], flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "probe_shift.json").write_text(json.dumps(res))
