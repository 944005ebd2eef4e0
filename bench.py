#!/usr/bin/env python
"""bench.py — tiles/sec of the segmentation training / inference hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # product arm (sm_100a kernels)
  python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference's CPU path (oracle port)

Headline workload (BASELINE.json configs[1]): UNet++-ResNet50, 4-band 512x512 synthetic uint8 tiles, 5 classes,
bf16 compute, batch 32 per GPU, train step = normalise + forward + CE loss + backward + (all-reduce) + Adam, replayed
from one CUDA graph at every N.  One JSON line on stdout (rank 0).  `value` = tiles/s with inputs resident in HBM;
`e2e` = the same step fed from pinned host memory (H2D inside the timed region) with the loss read back (D2H);
`roofline` = algorithmic FLOPs / measured kernel time of the forward + dgrad (and weight-gradient) convolutions;
`cpu_baseline` = the reference's CPU path on the box's cores; `library_baseline` = the reference's module stack in eager
bf16 autocast on the same GPU; `workloads` = BASELINE configs[2] SegFormer-B2, [3] DOFA-base + UperNet, [4] SegFormer-B5
sliding-window inference, each with the same keys (`--workloads headline` skips them, `--workload X` makes X the headline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "geo-deep-learning_b200"))

WORKLOADS = {
    # BASELINE.json configs[1] — the configuration the single-GPU metric is quoted on (default)
    "unetpp_r50": {"name": "unetpp_resnet50_4band_512_k5_b32", "family": "unetpp", "encoder": "resnet50", "bands": 4,
                   "tile": 512, "classes": 5, "batch_per_gpu": 32, "train_gflop_per_tile": 1380.7},
    # BASELINE.json configs[2] — SegFormer-B2, 3-band 512x512, batch 16 / GPU
    "segformer_b2": {"name": "segformer_b2_3band_512_k5_b16", "family": "segformer", "encoder": "mit_b2", "bands": 3,
                     "tile": 512, "classes": 5, "batch_per_gpu": 16, "train_gflop_per_tile": 363.1},
    # BASELINE.json configs[3] — DOFA-base (frozen, configs/dofa_config_RGB.yaml:57) + UperNet, 6-band 512x512.
    # GFLOP/tile = encoder forward 284.7 (12 ViT-B blocks at 1297 tokens + patch embed) + 3 x 443.05 (neck/UperNet/heads)
    "dofa_base": {"name": "dofa_base_upernet_6band_512_k5_b16", "family": "dofa", "encoder": "dofa_base", "bands": 6,
                  "tile": 512, "classes": 5, "batch_per_gpu": 16, "train_gflop_per_tile": 1613.8,
                  "wavelengths": [0.49, 0.56, 0.665, 0.842, 1.61, 2.19]},
    # the same model with the encoder trained too (3 x 666.1 = 1998.3 GFLOP / tile); eager launches: the weight
    # generator's backward is torch autograd, which a CUDA-graph capture cannot include
    "dofa_base_unfrozen": {"name": "dofa_base_upernet_6band_512_k5_b16_encoder_trained", "family": "dofa", "encoder": "dofa_base",
                           "bands": 6, "tile": 512, "classes": 5, "batch_per_gpu": 16, "train_gflop_per_tile": 1998.3,
                           "wavelengths": [0.49, 0.56, 0.665, 0.842, 1.61, 2.19], "unfrozen": True},
    # BASELINE.json configs[4] — inference only: SegFormer-B5, 4-band raster, 512-pixel windows at stride 256, windows dealt
    # round-robin to the ranks (strong scaling).  219.5 GFLOP per window (torch flop counter on the oracle's forward).
    "segformer_b5_infer": {"name": "segformer_b5_4band_sliding512_s256", "family": "infer", "encoder": "mit_b5", "bands": 4,
                           "tile": 512, "classes": 5, "batch_per_gpu": 16, "train_gflop_per_tile": 219.5, "raster": 10000},
}
WORKLOAD = WORKLOADS["unetpp_r50"]
TRAIN_GFLOP_PER_TILE = WORKLOAD["train_gflop_per_tile"]  # 3 x forward GFLOP (2 FLOP / MAC)
MEAN, STD = [0.5] * 6, [0.2] * 6


def _select_workload(key: str) -> None:
    global WORKLOAD, TRAIN_GFLOP_PER_TILE
    WORKLOAD = WORKLOADS[key]
    TRAIN_GFLOP_PER_TILE = WORKLOAD["train_gflop_per_tile"]


def _dram_traffic(family: str) -> tuple[float | None, str | None]:
    """DRAM bytes per launch of the forward/dgrad conv kernels (conv_fwd_kernel + conv3x3_rows_kernel), from the
    newest committed ncu pass (dram__bytes_read.sum + dram__bytes_write.sum per launch, tools/summarize_ncu_launches.py)."""
    import re
    best = None
    for f in sorted((ROOT / "profiles").glob(f"r*_dram_traffic_{family}.json")):
        m = re.search(r"r(\d+)_run(\d+)_", f.name)
        key = (int(m.group(1)), int(m.group(2))) if m else (0, 0)
        if best is None or key > best[0]:
            best = (key, f)
    if best is None:
        return None, None
    k = json.loads(best[1].read_text())["kernels"]
    rows = [k[n] for n in ("conv_fwd_kernel", "conv3x3_rows_kernel") if n in k]
    launches = sum(r["launches"] for r in rows)
    if not launches:
        return None, None
    total = sum((r["dram_read_GB"] + r["dram_write_GB"]) * 1e9 for r in rows)
    return total / launches, f"profiles/{best[1].name}"


def _peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        d["source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 2.0:  # first sample in hand before the timed region starts
                time.sleep(0.02)
        except OSError:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 8 and r[2].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =================================================================================================
# reference arm: the reference's own CPU path (oracle port of smp.UnetPlusPlus + torch CE + Adam)
# =================================================================================================
def cpu_reference_steps(steps: int, warmup: int, tiles_per_step: int = 1, budget_s: float | None = None) -> dict:
    import torch
    import torch.nn.functional as F
    from oracle import tensors as ot

    # all the threads torch will use on this box: its default = the cores this process may run on
    # (sched affinity / cgroup aware); forcing os.cpu_count() oversubscribes shared hosts and runs slower
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 (torch then reports 1 thread): use the physical core count instead — one
    # thread per LOGICAL cpu oversubscribes the SMT siblings and ran 30x slower on the GPU box (run 12)
    want = torch.get_num_threads()
    if want <= 1 or "OMP_NUM_THREADS" in os.environ:
        try:
            import psutil
            want = psutil.cpu_count(logical=False) or avail
        except Exception:  # noqa: BLE001
            want = max(1, avail // 2)
    cores = max(1, min(avail, want))
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    w = WORKLOAD
    nb = w["bands"]
    if w["family"] == "dofa":
        tiles_per_step = max(2, tiles_per_step)  # the PPM's 1x1 bin needs > 1 value per channel for train-mode BN
    g = torch.Generator().manual_seed(1234)
    raw = torch.randint(0, 256, (tiles_per_step, nb, w["tile"], w["tile"]), generator=g, dtype=torch.uint8)
    mask = torch.randint(0, w["classes"], (tiles_per_step, w["tile"] // 32, w["tile"] // 32), generator=g)
    mask = mask.repeat_interleave(32, 1).repeat_interleave(32, 2)
    mean, std = torch.tensor(MEAN[:nb]).view(-1, 1), torch.tensor(STD[:nb]).view(-1, 1)
    if w["family"] == "unetpp":
        from oracle.unetpp import UnetPlusPlusOracle
        model = UnetPlusPlusOracle(w["encoder"], nb, w["classes"]).train()
        params = list(model.parameters())

        def fwd(x):
            return model(x)
        what = "oracle port of smp.UnetPlusPlus: smp itself is not installable offline"
    elif w["family"] == "infer":
        return cpu_reference_infer(steps, warmup, budget_s, cores)
    elif w["family"] == "dofa":
        from oracle import dofa as od, upernet as ou
        enc_sd = od.init_state_dict(768, 12, w["tile"])
        sd = ou.init_state_dict(768, 256, w["classes"])
        sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        params = [v for v in sd.values() if v.requires_grad]
        wl = torch.tensor(w["wavelengths"])
        unfrozen = bool(w.get("unfrozen"))
        if unfrozen:
            enc_sd = {k: (v.requires_grad_(True) if v.is_floating_point() and k != "pos_embed" else v) for k, v in enc_sd.items()}
            params += [v for v in enc_sd.values() if v.requires_grad]

        def fwd(x):
            with torch.set_grad_enabled(unfrozen):  # frozen encoder (freeze_layers: ["encoder"]) unless the workload trains it
                feats = od.dofa_forward(enc_sd, x, wl)
            return ou.upernet_forward(sd, feats, (w["tile"], w["tile"]), training=True)
        what = ("functional restatement of the reference's DOFASegmentationModel (frozen DOFAv2 + UperNet), pinned to it "
                "by golden vectors")
    else:
        from oracle import segformer as osf
        sd = osf.init_state_dict(w["encoder"], nb, w["classes"])
        sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        params = [v for v in sd.values() if v.requires_grad]

        def fwd(x):
            return osf.segformer_forward(sd, x, w["encoder"], training=True)
        what = "functional restatement of the reference's SegFormerSegmentationModel, pinned to it by golden vectors"
    opt = torch.optim.Adam(params, lr=1e-4)

    def step() -> float:
        x = ot.standardization(ot.normalization(raw.float()), mean, std)
        opt.zero_grad(set_to_none=True)
        out = fwd(x)
        loss = (F.cross_entropy(out[0], mask) + 0.4 * F.cross_entropy(out[1], mask)) if isinstance(out, tuple) \
            else F.cross_entropy(out, mask)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    return {"value": tiles_per_step * len(times) / total, "unit": "tiles/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} step(s) x {tiles_per_step} tile(s) of {w['name']} (fwd+CE+bwd+Adam, fp32, "
                      f"{cores} threads, {what})",
            "ms_per_step": 1e3 * total / len(times), "steps_timed": len(times), "tiles_per_step": tiles_per_step}


def cpu_reference_infer(steps: int, warmup: int, budget_s: float | None, cores: int) -> dict:
    """Reference eval path of one window on the host: normalise -> SegFormer forward -> softmax/argmax."""
    import torch
    from oracle import segformer as osf
    from oracle import tensors as ot
    w = WORKLOAD
    nb = w["bands"]
    sd = osf.init_state_dict(w["encoder"], nb, w["classes"])
    g = torch.Generator().manual_seed(1234)
    raw = torch.randint(0, 256, (1, nb, w["tile"], w["tile"]), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(MEAN[:nb]).view(-1, 1), torch.tensor(STD[:nb]).view(-1, 1)

    def step() -> None:
        with torch.no_grad():
            x = ot.standardization(ot.normalization(raw.float()), mean, std)
            osf.segformer_forward(sd, x, w["encoder"], training=False).softmax(dim=1).argmax(dim=1)

    for _ in range(warmup):
        step()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    return {"value": len(times) / total, "unit": "tiles/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} window(s) of {w['name']} (normalise + forward + softmax/argmax, fp32, {cores} threads, "
                      "functional restatement of the reference's SegFormerSegmentationModel, pinned to it by golden vectors)",
            "ms_per_step": 1e3 * total / len(times), "steps_timed": len(times), "tiles_per_step": 1}


def main_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup, 1)
    infer = WORKLOAD["family"] == "infer"
    line = {
        "impl": "reference", "metric": "512x512 multi-band tiles/sec (inference fwd, sliding window)" if infer
        else "512x512 multi-band tiles/sec (train fwd+bwd)", "value": r["value"],
        "unit": "tiles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if infer else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "tiles_per_step": r["tiles_per_step"],
                   "note": "reference CPU path, bounded sample"},
        "cpu_baseline": {"value": r["value"], "unit": "tiles/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# =================================================================================================
# library baseline: the reference's own modules (oracle restatements, pinned to them) as a user of the reference runs
# them on this GPU — stock PyTorch eager, torch.autocast(bf16), cuDNN / cuBLAS kernels, torch.optim.Adam(fused=True).
# BASELINE.md §3 calls this "the real bar"; it is reported beside the product arm, never used by it.
# =================================================================================================
def library_baseline_steps(w: dict, dev, batch: int, steps: int = 5, warmup: int = 2) -> dict:
    import torch
    import torch.nn.functional as F

    torch.backends.cudnn.benchmark = True
    nb, T, K = w["bands"], w["tile"], w["classes"]
    g = torch.Generator().manual_seed(4321)
    infer = w["family"] == "infer"
    # what the reference's DataLoader hands over: a float32 NCHW batch, already normalised + standardised (resident in HBM)
    x = torch.randn(batch, nb, T, T, generator=g).to(dev)
    mask = torch.randint(0, K, (batch, T // 32, T // 32), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).to(dev)
    torch.manual_seed(0)
    if w["family"] == "unetpp":
        from oracle.unetpp import UnetPlusPlusOracle
        model = UnetPlusPlusOracle(w["encoder"], nb, K).to(dev).train()
        params = list(model.parameters())

        def fwd(inp):
            return model(inp)
    elif w["family"] == "dofa":
        from oracle import dofa as od, upernet as ou
        enc_sd = {k: v.to(dev) for k, v in od.init_state_dict(768, 12, T).items()}
        sd = {k: v.to(dev) for k, v in ou.init_state_dict(768, 256, K).items()}
        sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        params = [v for v in sd.values() if v.requires_grad]
        wl = torch.tensor(w["wavelengths"], device=dev)

        def fwd(inp):
            with torch.no_grad():
                feats = od.dofa_forward(enc_sd, inp, wl)
            return ou.upernet_forward(sd, feats, (T, T), training=True)
    else:
        from oracle import segformer as osf
        sd = {k: v.to(dev) for k, v in osf.init_state_dict(w["encoder"], nb, K).items()}
        if not infer:
            sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        params = [v for v in sd.values() if v.requires_grad]

        def fwd(inp):
            return osf.segformer_forward(sd, inp, w["encoder"], training=not infer)
    opt = torch.optim.Adam(params, lr=1e-4, fused=True) if params else None

    def step() -> None:
        if infer:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                fwd(x).softmax(dim=1).argmax(dim=1)
            return
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = fwd(x)
        loss = (F.cross_entropy(out[0].float(), mask) + 0.4 * F.cross_entropy(out[1].float(), mask)) if isinstance(out, tuple) \
            else F.cross_entropy(out.float(), mask)
        loss.backward()
        if w["family"] in ("segformer", "dofa"):
            torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": batch / (ms / 1e3), "unit": "tiles/s", "ms_per_step": ms, "batch": batch, "steps": steps,
           "what": ("the reference's module stack (oracle restatement pinned to it) on the SAME GPU: stock PyTorch eager, "
                    "torch.autocast(bf16), cuDNN benchmark mode, " + ("no_grad forward + softmax/argmax" if infer else
                    "fwd + CE + bwd + fused Adam") + "; inputs resident in HBM as float32 NCHW")}
    del opt, params
    return out


# =================================================================================================
# product arm
# =================================================================================================
def main_product(args) -> None:
    import torch
    import torch.distributed as dist

    from gdl_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    ctx = (world, rank, local, dev)

    head_key = args.workload
    line = _run_workload(args, head_key, ctx, steps=args.steps, warmup=args.warmup, headline=True)
    # the other BASELINE.json configurations ride along in the same line: configs[2]
    # SegFormer-B2 training, configs[3] DOFA-base + UperNet training, configs[4] SegFormer-B5 sliding-window inference
    extra = {}
    if args.workloads == "all" and head_key == "unetpp_r50":
        for key in ("segformer_b2", "dofa_base", "segformer_b5_infer"):
            sub_steps = max(1, min(args.steps, 2 if key == "segformer_b5_infer" else 6))
            try:
                sub = _run_workload(args, key, ctx, steps=sub_steps, warmup=min(args.warmup, 3) if key != "segformer_b5_infer" else 1,
                                    headline=False)
            except Exception as e:  # noqa: BLE001  (a ride-along workload must not take the headline's line with it)
                import traceback
                traceback.print_exc()
                sub = {"error": f"{type(e).__name__}: {e}"[:300]} if rank == 0 else None
            if sub is not None:
                extra[key] = sub
    if rank == 0:
        if extra:
            line["workloads"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _run_workload(args, key: str, ctx, *, steps: int, warmup: int, headline: bool):
    import gc

    import torch
    _select_workload(key)
    try:
        if WORKLOAD["family"] == "infer":
            return _run_infer(args, ctx, steps, warmup, headline)
        return _run_train(args, ctx, steps, warmup, headline)
    finally:
        gc.collect()
        torch.cuda.empty_cache()
        _select_workload(args.workload)


def _run_train(args, ctx, steps: int, warmup: int, headline: bool):
    import torch
    import torch.distributed as dist

    from gdl_b200 import ops
    from gdl_b200.models.unetpp import UnetPlusPlus
    from gdl_b200.trainer import FusedTrainer

    world, rank, local, dev = ctx
    w = WORKLOAD
    B, C, T, K = (args.batch if headline and args.batch else w["batch_per_gpu"]), w["bands"], w["tile"], w["classes"]
    torch.manual_seed(0)  # identical initial weights on every rank (what DDP's broadcast gives)
    cuda_graph = args.cuda_graph
    if w["family"] == "unetpp":
        model = UnetPlusPlus(w["encoder"], in_channels=C, classes=K, compute_dtype=torch.bfloat16).to(dev).train()
    elif w["family"] == "dofa":
        from gdl_b200.models.dofa import DOFASegmentationModel
        model = DOFASegmentationModel(w["encoder"], (T, T), None if w.get("unfrozen") else ["encoder"], K,
                                      compute_dtype=torch.bfloat16).to(dev).train()
        if w.get("unfrozen"):
            cuda_graph = 0
        model.wavelengths = torch.tensor(w["wavelengths"], device=dev)
    else:
        from gdl_b200.models.segformer import SegFormer
        model = SegFormer(w["encoder"], in_channels=C, num_classes=K, compute_dtype=torch.bfloat16).to(dev).train()
    use_graph = bool(cuda_graph) and (world == 1 or cuda_graph >= 2)
    trainer = FusedTrainer(model, ops.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-4, mean=MEAN[:C], std=STD[:C],
                           image_max=255.0, sync_bn=bool(args.sync_bn),
                           clip_grad_norm=1.0 if w["family"] in ("segformer", "dofa") else None, cuda_graph=use_graph)

    # synthetic tiles: NBUF distinct batches so consecutive steps never re-read the same input (and the
    # per-step working set, tens of GB of activations, is far larger than the 126 MB L2 anyway)
    NBUF = 4
    g = torch.Generator().manual_seed(1234 + rank)
    host_img = [torch.randint(0, 256, (B, T, T, C), generator=g, dtype=torch.uint8).pin_memory() for _ in range(NBUF)]
    host_msk = []
    for _ in range(NBUF):
        m = torch.randint(0, K, (B, T // 32, T // 32), generator=g, dtype=torch.uint8)
        host_msk.append(m.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous().pin_memory())
    dev_img = [t.to(dev) for t in host_img]
    dev_msk = [t.to(dev) for t in host_msk]

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps: int) -> float:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nsteps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident(i: int):
        return trainer.step(dev_img[i % NBUF], dev_msk[i % NBUF])

    loss_host = torch.zeros(1).pin_memory()

    def step_e2e(i: int):
        img = host_img[i % NBUF].to(dev, non_blocking=True)
        msk = host_msk[i % NBUF].to(dev, non_blocking=True)
        loss = trainer.step(img, msk)
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller gets the loss value every step
        return loss_host

    for i in range(max(warmup, 2)):  # >= 2: the second step captures the graph
        step_resident(i)
    if not headline:
        # ride-along workloads: enough steps for ~1.5 s of timed work (nvidia-smi samples clocks every 100 ms); the
        # headline times EXACTLY the --steps the driver asked for
        t_one = timed(step_resident, 1)
        steps = int(max(steps, min(60, 1500.0 / max(t_one, 1.0) + 1)))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(step_resident, steps)
    launches = trainer.launches_per_step * steps  # kernels of libgdlb200.so per step (counted at capture)
    clk = clocks.stop() if rank == 0 else None

    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, steps)

    # ---- roofline of the dominant kernels: per-launch CUDA events around every tensor-core conv launch
    # (eager launches: events cannot be read back from inside a replayed graph)
    trainer.cuda_graph = False
    prof = ops.ConvProfiler()
    ops.set_conv_profiler(prof)
    nprof = max(1, min(3, steps)) if headline else 1
    for i in range(nprof):
        step_resident(i)
    torch.cuda.synchronize()
    ops.set_conv_profiler(None)
    roof = prof.summary(nprof)
    if headline and args.table and rank == 0:
        prof.dump_table(args.table, nprof)
    deterministic = ops.deterministic()
    bn_exchange_kind = trainer.bn_exchange_kind
    del trainer, model, dev_img, dev_msk
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    cpu = lib = None
    if rank == 0 and headline and not args.no_cpu_baseline:
        cpu = cpu_reference_steps(steps=3, warmup=1, tiles_per_step=1, budget_s=20.0)
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            lib = library_baseline_steps(w, dev, B, steps=5 if headline else 3, warmup=2)
        except Exception as e:  # noqa: BLE001  (a baseline that cannot run must not take the product's line with it)
            lib = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        return None
    peaks = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    tiles = B * world * steps
    value = tiles / (ms / 1e3)
    fwd = roof["conv_fwd_kernel"]
    traffic, traffic_src = _dram_traffic({"unetpp": "unetpp", "segformer": "segformer", "dofa": "dofa"}[w["family"]])
    if w["family"] == "segformer" and ops.option("decoder_folded"):
        # the committed SegFormer ncu pass predates the folded decoder (its launch list contains the 4*emb -> emb fuse
        # product): no per-launch DRAM figure for the launches that run now
        traffic, traffic_src = None, None
    line = {
        "metric": "512x512 multi-band tiles/sec (train fwd+bwd)", "value": value, "unit": "tiles/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": w["name"], "global_batch": B * world, "parallelism": f"dp{world}",
                   "loss": "cross_entropy", "optimizer": "adam", "sync_bn": bool(args.sync_bn and world > 1), "cuda_graph": use_graph,
                   "sra_fused": bool(ops.option("sra_fused")), "mha_flash": bool(ops.option("mha_flash")),
                   "deterministic_reductions": deterministic, "syncbn_exchange": bn_exchange_kind,
                   "pdl": bool(ops.option("pdl")),
                   **({"decoder_folded": bool(ops.option("decoder_folded"))} if w["family"] == "segformer" else {}),
                   "l2": f"{NBUF} rotating input batches; per-step working set >> 126 MB L2"},
        "clocks": clk,
        "e2e": {"value": tiles / (ms_e2e / 1e3), "unit": "tiles/s",
                "h2d_bytes_per_step": B * T * T * C + B * T * T, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": launches,
        "model_tflops": TRAIN_GFLOP_PER_TILE * value / world / 1e3,
        **({"flops_note": "model_tflops and roofline.whole_step count the reference topology's 363.1 GFLOP per tile; with "
            "decoder_folded the 4*emb -> emb fuse product (3 x 77.3 GFLOP per tile) runs in front of the resizes at each level's "
            "own resolution and 131.2 GFLOP per tile are executed"} if w["family"] == "segformer" and ops.option("decoder_folded") else {}),
        "roofline": {"bound": "tensor", "kernel": "conv_fwd_kernel + conv3x3_rows_kernel (forward + dgrad launches)",
                     "achieved": fwd["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": fwd["tflops"] / peak_tf, "flops": "algorithmic (pixel-packed / channel-padded launches count their unpadded shapes)",
                     "traffic": traffic, "traffic_unit": "DRAM bytes per launch",
                     "traffic_source": traffic_src, "peak_source":
                     f"{peaks['source']} bf16_tflops_sustained", "launches_per_step": fwd["launches_per_step"],
                     "share_of_step": fwd["ms_per_step"] / (ms / steps),
                     "wgrad": {"kernel": "conv_wgrad_kernel + wgrad3x3_rows_kernel (+ their ordered-reduce passes)",
                               "achieved": roof["conv_wgrad_kernel"]["tflops"],
                               "frac": roof["conv_wgrad_kernel"]["tflops"] / peak_tf,
                               "share_of_step": roof["conv_wgrad_kernel"]["ms_per_step"] / (ms / steps)},
                     "whole_step": {"achieved": TRAIN_GFLOP_PER_TILE * value / world / 1e3, "frac":
                                    TRAIN_GFLOP_PER_TILE * value / world / 1e3 / peak_tf}},
        "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
        "library_baseline": lib,
    }
    return line


def _run_infer(args, ctx, steps: int, warmup: int, headline: bool):
    """Sliding-window inference over one synthetic raster, windows dealt round-robin to the ranks (strong scaling)."""
    import torch
    import torch.distributed as dist

    from gdl_b200 import ops
    from gdl_b200.inference import SlidingWindowSegmenter, window_origins
    from gdl_b200.models.segformer import SegFormer

    world, rank, local, dev = ctx
    w = WORKLOAD
    B, C, T, K = (args.batch if headline and args.batch else w["batch_per_gpu"]), w["bands"], w["tile"], w["classes"]
    R = args.raster or w["raster"]
    torch.manual_seed(0)
    model = SegFormer(w["encoder"], in_channels=C, num_classes=K, compute_dtype=torch.bfloat16).to(dev).eval()
    # the window-batch forward (normalise .. logits) is replayed from a CUDA graph by default (--cuda-graph >= 2; validated by
    # tests/test_zz1_inference_gpu.py, +4.6 % on B200: profiles/r02_run17_bench_infer_cg{2,3}.json); 0 / 1 = eager launches
    seg = SlidingWindowSegmenter(model, tile=T, stride=T // 2, batch=B, mean=MEAN[:C], std=STD[:C],
                                 cuda_graph=args.cuda_graph >= 2)
    nwin = len(window_origins(R, T, T // 2)) ** 2
    g = torch.Generator().manual_seed(1234)  # the same raster on every rank
    host = torch.randint(0, 256, (R, R, C), generator=g, dtype=torch.uint8).pin_memory()
    resident = host.to(dev)
    out_host = torch.empty((R, R), dtype=torch.uint8).pin_memory()

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps: int) -> float:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nsteps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident(i: int):
        return seg.predict(resident)

    def step_e2e(i: int):
        out_host.copy_(seg.predict(host), non_blocking=True)  # raster H2D and class map D2H inside the timed region
        torch.cuda.current_stream().synchronize()

    n0 = ops.launch_count()
    for i in range(max(1, warmup)):
        step_resident(i)
    launches_per_step = (ops.launch_count() - n0) // max(1, warmup)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(step_resident, steps)
    clk = clocks.stop() if rank == 0 else None
    step_e2e(0)
    ms_e2e = timed(step_e2e, steps)
    prof = ops.ConvProfiler()
    ops.set_conv_profiler(prof)
    step_resident(0)
    torch.cuda.synchronize()
    ops.set_conv_profiler(None)
    roof = prof.summary(1)
    del seg, model, resident
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    cpu = lib = None
    if rank == 0 and headline and not args.no_cpu_baseline:
        cpu = cpu_reference_steps(steps=3, warmup=1, tiles_per_step=1, budget_s=20.0)
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            lib = library_baseline_steps(w, dev, B, steps=3, warmup=2)
        except Exception as e:  # noqa: BLE001
            lib = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if rank != 0:
        return None
    peaks = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    value = nwin * steps / (ms / 1e3)
    fwd = roof["conv_fwd_kernel"]
    line = {
        "metric": "512x512 multi-band tiles/sec (inference fwd, sliding window)", "value": value, "unit": "tiles/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": w["name"], "raster": [R, R, C], "windows": nwin, "window_batch": B,
                   "parallelism": f"windows round-robin over {world} rank(s) + one all-reduce of the logit sums",
                   "cuda_graph": args.cuda_graph >= 2, "sra_fused": bool(ops.option("sra_fused")),
                   "pdl": bool(ops.option("pdl")), "decoder_folded": bool(ops.option("decoder_folded")),
                   "l2": f"raster {R * R * C / 1e6:.0f} MB and activations >> 126 MB L2"},
        "clocks": clk,
        "e2e": {"value": nwin * steps / (ms_e2e / 1e3), "unit": "tiles/s", "h2d_bytes_per_step": R * R * C,
                "d2h_bytes_per_step": R * R, "ms_per_step": ms_e2e / steps},
        "gpu_launches": launches_per_step * steps,
        "model_tflops": w["train_gflop_per_tile"] * value / world / 1e3,
        "roofline": {"bound": "tensor", "kernel": "conv_fwd_kernel (every GEMM / conv of the forward)",
                     "achieved": fwd["tflops"], "peak": peak_tf, "unit": "TFLOP/s", "frac": fwd["tflops"] / peak_tf,
                     "traffic": None, "peak_source": f"{peaks['source']} bf16_tflops_sustained",
                     "launches_per_step": fwd["launches_per_step"],
                     "share_of_step": fwd["ms_per_step"] / (ms / steps)},
        "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
        "library_baseline": lib,
    }
    return line


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="tiles per GPU (default: the workload's 32)")
    ap.add_argument("--sync-bn", type=int, default=1, help="SyncBatchNorm statistics when N > 1 (reference YAMLs: true)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=2,
                    help="0: eager launches; 1: capture the whole step in a CUDA graph at N=1 only; 2 (default): also at N>1, the NCCL "
                         "collectives (SyncBN sums, flat-gradient all-reduce) captured with the kernels (validated at N=2: "
                         "profiles/r02_run6_*)")
    ap.add_argument("--raster", type=int, default=0, help="segformer_b5_infer: raster side in pixels (default 10000)")
    ap.add_argument("--table", default="", help="write the per-launch conv profile (shape, ms, TFLOP/s) to this JSON file")
    ap.add_argument("--workload", default="unetpp_r50", choices=sorted(WORKLOADS))
    ap.add_argument("--workloads", default="all", choices=["all", "headline"],
                    help="all: after the headline (default workload only) also run BASELINE configs[2..4] and attach them as `workloads`")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip the stock-PyTorch eager bf16 run of the reference's modules on the same GPU")
    args = ap.parse_args()
    _select_workload(args.workload)
    if args.impl == "reference":
        main_reference(args)
    else:
        main_product(args)


if __name__ == "__main__":
    main()
