"""TEST INFRASTRUCTURE: run the GPU test files on the CPU against the HOST-COMPILED scalar CUDA kernels.

`install(monkeypatch)` points gdl_b200's ctypes binding at tests/hostemu/_build/libgdlb200_hostemu.so (the product's own .cu
sources compiled for the host, see build.py / cuda_hostemu.h) and lets the wrappers accept CPU tensors.  The tensor-core
entry points (tcgen05 / TMA: conv forward / wgrad and the weight packing around them) cannot run on a CPU and are taken
from the torch emulation of tests/cpu_kernel_emulation.py instead.  Nothing here is imported by the product.
"""
from __future__ import annotations

import ctypes as C
import inspect
import itertools
import os
import re
import sys
import types
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
TESTS = HERE.parent

TENSOR_CORE_OPS = ("conv2d_fwd", "conv2d_wgrad", "pack_conv_weight", "unpack_conv_wgrad", "widen_conv_weight", "fold_widened_wgrad")

_lib = None
_ws = None  # host copy of the reduction workspace (gdl_set_workspace): the ordered-reduction paths run on the CPU too


def load_lib():
    global _lib
    if _lib is None:
        from . import build
        from gdl_b200 import _lib as L
        lib = C.CDLL(str(build.build()), mode=os.RTLD_NOW)
        lib.gdl_last_error.restype = C.c_char_p
        lib.gdl_version.restype = C.c_int
        for name, args in L._SIGS.items():
            fn = getattr(lib, name, None)
            if fn is not None:
                fn.argtypes = args
                fn.restype = C.c_int
        _lib = lib
    return _lib


def install(monkeypatch, torch_convs: bool = False, async_seed: int | None = None) -> None:
    """torch_convs: take the tensor-core entry points from the torch emulation instead of running the tcgen05 / TMA kernels on
    the functional model of hostemu_tc.cpp (faster for whole-model steps).
    async_seed: None = GDL_HOSTEMU_ASYNC from the environment (unset: synchronous completion); otherwise the asynchronous mode
    of the model with that seed — TMA copies / MMAs / commits complete at random later points (even seed: operands read at issue,
    odd: at completion), which exposes missing waits and premature reuse of stages, accumulators and staging tiles."""
    import cpu_kernel_emulation as emu
    from gdl_b200 import _lib as L
    from gdl_b200 import ops
    lib = load_lib()
    if async_seed is None:
        async_seed = int(os.environ.get("GDL_HOSTEMU_ASYNC", "-1") or -1)
    lib.hostemu_set_async(int(async_seed))
    monkeypatch.setattr(L, "_lib", lib)
    monkeypatch.setattr(L, "stream_ptr", lambda: C.c_void_p(0))
    monkeypatch.setattr(L, "ptr", lambda t: C.c_void_p(0 if t is None else t.data_ptr()))
    monkeypatch.setattr(ops, "require_cuda", lambda t, what: None)
    global _ws
    if _ws is None and os.environ.get("GDL_HOSTEMU_DET", "1") != "0":
        lib.gdl_query_workspace_bytes.restype = C.c_longlong
        n = int(lib.gdl_query_workspace_bytes())
        _ws = torch.zeros(n + 512, dtype=torch.uint8)
        off = (-_ws.data_ptr()) % 256
        _ws = _ws[off:off + n]
        L.check(lib.gdl_set_workspace(C.c_void_p(_ws.data_ptr()), n, C.c_void_p(0)))
    monkeypatch.setattr(L, "workspace_tensor", lambda: _ws)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    if torch_convs:
        for n in TENSOR_CORE_OPS:
            monkeypatch.setattr(ops, n, getattr(emu, n))
    emu.set_work_dtype(torch.float32)


def _rewrite(src: str) -> str:
    return (src.replace('device="cuda"', 'device="cpu"').replace(".cuda()", ".cpu()").replace('torch.autocast("cuda"', 'torch.autocast("cpu"')
            .replace(".pin_memory()", "").replace(".is_cuda", ".is_cpu").replace("pytestmark = pytest.mark.gpu", "pytestmark = []"))


def load_test_module(name: str) -> types.ModuleType:
    """a GPU test file re-targeted at CPU tensors (module name prefixed so that it never shadows the real one)"""
    key = f"hostemu_{name}"
    if key in sys.modules:
        return sys.modules[key]
    path = TESTS / f"{name}.py"
    mod = types.ModuleType(key)
    mod.__file__ = str(path)
    sys.modules[key] = mod
    src = _rewrite(path.read_text())
    for dep in set(re.findall(r"^\s*from (test_\w+) import", src, flags=re.M)):  # helper imports between GPU test files
        load_test_module(dep)
        src = re.sub(rf"from {dep} import", f"from hostemu_{dep} import", src)
    exec(compile(src, str(path), "exec"), mod.__dict__)
    return mod


def cases(name: str, include=None, exclude=()):
    """(function name, kwargs, id) for every parametrisation of the module's tests"""
    mod = load_test_module(name)
    out = []
    for fname, fn in vars(mod).items():
        if not fname.startswith("test_") or not callable(fn):
            continue
        if include is not None and fname not in include or fname in exclude:
            continue
        axes = []
        for mark in getattr(fn, "pytestmark", []):
            if mark.name == "parametrize":
                names = [a.strip() for a in mark.args[0].split(",")]
                axes.append([dict(zip(names, v if isinstance(v, (tuple, list)) and len(names) > 1 else (v,))) for v in mark.args[1]])
        for combo in itertools.product(*axes) if axes else [()]:
            kw = {}
            for d in combo:
                kw.update(d)
            ident = f"{name}::{fname}" + ("[" + "-".join(str(v).replace("torch.", "") for v in kw.values()) + "]" if kw else "")
            out.append((fname, kw, ident))
    return out


def run_case(name: str, fname: str, kw: dict, tmp_path=None) -> None:
    fn = getattr(load_test_module(name), fname)
    kwargs = dict(kw)
    sig = inspect.signature(fn)
    if "cuda" in sig.parameters:
        kwargs["cuda"] = torch.device("cpu")
    if "tmp_path" in sig.parameters:
        kwargs["tmp_path"] = tmp_path
    # module-level pytest fixtures (option switches with a teardown) are driven by hand
    mod, teardown = load_test_module(name), []
    for pname, par in sig.parameters.items():
        if pname in kwargs or par.default is not inspect.Parameter.empty:
            continue
        fx = getattr(mod, pname, None)
        raw = getattr(fx, "_get_wrapped_function", None)
        raw = raw() if raw else getattr(fx, "__wrapped__", None)
        if raw is None:
            raise TypeError(f"{fname}: no value for parameter {pname}")
        val = raw()
        if inspect.isgenerator(val):
            teardown.append(val)
            val = next(val)
        kwargs[pname] = val
    try:
        fn(**kwargs)
    finally:
        for gen in teardown:
            for _ in gen:
                pass
