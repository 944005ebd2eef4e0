#!/usr/bin/env python
"""Dry run of GPU test files on the CPU (build container has no GPU): the test source is rewritten to CPU devices and the
kernel wrappers are replaced by the torch emulation of tests/cpu_kernel_emulation.py.  This checks the TESTS THEMSELVES
(shapes, keyword arguments, call order, tolerances that do not depend on 16-bit rounding) before they meet a B200 — it
says nothing about the CUDA kernels.  Expected artefacts of the emulation: `pytest.raises(ValueError)` blocks that rely on the
real wrappers' argument checks, and assertions about CUDA-graph capture, fail here.

  python tools/dryrun_gpu_tests_on_cpu.py tests/test_zz2_augment_metrics_gpu.py [more files]
"""
from __future__ import annotations

import inspect
import sys
import traceback
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "geo-deep-learning_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))

import pytest  # noqa: E402
import torch  # noqa: E402

import cpu_kernel_emulation as emu  # noqa: E402


def _rewrite(src: str) -> str:
    return (src.replace('device="cuda"', 'device="cpu"').replace(".cuda()", ".cpu()").replace('torch.autocast("cuda"', 'torch.autocast("cpu"')
            .replace(".pin_memory()", "").replace(".is_cuda", ".is_cpu").replace("pytestmark = pytest.mark.gpu", "pytestmark = []"))


def load(path: Path) -> types.ModuleType:
    mod = types.ModuleType(path.stem)
    mod.__file__ = str(path)
    sys.modules[path.stem] = mod
    exec(compile(_rewrite(path.read_text()), str(path), "exec"), mod.__dict__)
    return mod


def main() -> int:
    emu.install_global()
    failed = 0
    for arg in sys.argv[1:]:
        path = Path(arg).resolve()
        for dep in ("test_segformer_gpu", "test_wds_feeder_cpu"):  # helper modules the files import from
            if dep not in sys.modules and (ROOT / "tests" / f"{dep}.py").exists() and dep in path.read_text():
                load(ROOT / "tests" / f"{dep}.py")
        mod = load(path)
        for name, fn in list(vars(mod).items()):
            if not name.startswith("test_") or not callable(fn):
                continue
            params = [None]
            for mark in getattr(fn, "pytestmark", []):
                if mark.name == "parametrize":
                    names = [a.strip() for a in mark.args[0].split(",")]
                    params = [dict(zip(names, v if isinstance(v, (tuple, list)) and len(names) > 1 else (v,))) for v in mark.args[1]]
            for prm in params:
                kwargs = dict(prm or {})
                sig = inspect.signature(fn)
                if "cuda" in sig.parameters:
                    kwargs["cuda"] = torch.device("cpu")
                if "tmp_path" in sig.parameters:
                    import tempfile
                    kwargs["tmp_path"] = Path(tempfile.mkdtemp())
                label = f"{path.name}::{name}{prm or ''}"
                try:
                    fn(**kwargs)
                    print(f"ok      {label}")
                except pytest.skip.Exception as e:
                    print(f"skipped {label}: {e}")
                except Exception:  # noqa: BLE001
                    failed += 1
                    print(f"FAILED  {label}")
                    traceback.print_exc(limit=6)
    return 1 if failed else 0


if __name__ == "__main__":
    raise SystemExit(main())
