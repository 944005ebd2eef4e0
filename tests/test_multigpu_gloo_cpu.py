"""CPU, world_size = 2 over gloo: the N>1 host logic of the fused trainer (flat-gradient all-reduce,
SyncBatchNorm statistics exchange, identical updates on every rank).  Kernels are the float64 torch
emulation (tests/cpu_kernel_emulation.py); the property checked is the one DDP + SyncBN guarantee:
two ranks with B tiles each end up with exactly the parameters of one process training on the 2B tiles."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(k=4):
    g = torch.Generator().manual_seed(3)
    t = torch.randint(0, k, (4, 2, 2), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    raw = (t.unsqueeze(-1) * 60 + torch.randint(0, 20, (4, 32, 32, 3), generator=g)).to(torch.uint8)
    return raw, t


def _make_trainer(sync_bn, monkeypatch=None):
    import cpu_kernel_emulation as emu
    from gdl_b200.models.unetpp import UnetPlusPlus
    from gdl_b200.trainer import FusedTrainer
    if monkeypatch is None:
        emu.install_global()  # spawned worker: the process ends with the test
    else:
        emu.install(monkeypatch)  # pytest's own process: undone at teardown
    emu.set_work_dtype(torch.float64)
    torch.manual_seed(0)
    model = UnetPlusPlus("resnet18", in_channels=3, classes=4, compute_dtype=torch.float64).double().train()
    return FusedTrainer(model, emu.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-2, mean=[0.5] * 3, std=[0.25] * 3,
                        sync_bn=sync_bn, acc_dtype=torch.float64)


def _worker(rank, world, port, out):
    for p in (ROOT, ROOT / "geo-deep-learning_b200", ROOT / "tests"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)  # two workers on one host: full-width OpenMP teams spin against each other
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr = _make_trainer(sync_bn=True)
    assert tr.world == 2 and tr.sync_bn
    raw, t = _data()
    lo, hi = rank * 2, rank * 2 + 2  # rank-sharded tiles (weak scaling: 2 tiles per rank)
    losses = [float(tr.step(raw[lo:hi], t[lo:hi])) for _ in range(2)]
    torch.save({"flat": tr.flat.clone(), "losses": losses,
                "rm": tr.model.encoder.bn1.running_mean.clone()}, f"{out}/rank{rank}.pt")
    dist.destroy_process_group()


@pytest.fixture
def f64_work_dtype():
    import cpu_kernel_emulation as emu
    yield
    emu.set_work_dtype(torch.float32)


def test_two_ranks_equal_one_process_on_the_union(tmp_path, monkeypatch, f64_work_dtype):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["flat"], r1["flat"])  # every rank applies the same update
    assert torch.equal(r0["rm"], r1["rm"])
    # single process, all 4 tiles
    sys.path.insert(0, str(ROOT / "tests"))
    tr = _make_trainer(sync_bn=False, monkeypatch=monkeypatch)
    raw, t = _data()
    losses = [float(tr.step(raw, t)) for _ in range(2)]
    assert torch.allclose(tr.flat, r0["flat"], atol=1e-9, rtol=1e-7)
    assert torch.allclose(tr.model.encoder.bn1.running_mean, r0["rm"], atol=1e-10)
    # the global loss is the mean of the per-rank losses
    for i in range(2):
        assert abs(losses[i] - 0.5 * (r0["losses"][i] + r1["losses"][i])) < 1e-9


# ---------------------------------------------------------------------------------------------
# models that own their backward (SegFormer: `model.backward`, no per-closure progress hook): the bucketed all-reduce must
# still reduce EVERY bucket in EVERY step — three steps, so a bucket cursor left at the end of step 1 would show
# ---------------------------------------------------------------------------------------------
def _segformer_trainer(monkeypatch=None):
    import cpu_kernel_emulation as emu
    from gdl_b200.models.segformer import SegFormer
    from gdl_b200.trainer import FusedTrainer
    if monkeypatch is None:
        emu.install_global()
    else:
        emu.install(monkeypatch)
    emu.set_work_dtype(torch.float64)
    torch.manual_seed(0)
    model = SegFormer("mit_b0", in_channels=3, num_classes=4, compute_dtype=torch.float64).double().train()
    return FusedTrainer(model, emu.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.25] * 3,
                        sync_bn=True, acc_dtype=torch.float64)


def _segformer_data():
    g = torch.Generator().manual_seed(5)
    t = torch.randint(0, 4, (4, 2, 2), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    raw = (t.unsqueeze(-1) * 60 + torch.randint(0, 20, (4, 32, 32, 3), generator=g)).to(torch.uint8)
    return raw, t


def _segformer_worker(rank, world, port, out):
    for p in (ROOT, ROOT / "geo-deep-learning_b200", ROOT / "tests"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)  # two workers on one host: full-width OpenMP teams spin against each other
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr = _segformer_trainer()
    assert tr.world == 2 and tr.overlap_allreduce and len(tr._buckets) >= 2
    raw, t = _segformer_data()
    lo, hi = rank * 2, rank * 2 + 2
    losses = [float(tr.step(raw[lo:hi], t[lo:hi])) for _ in range(3)]
    torch.save({"flat": tr.flat.clone(), "losses": losses}, f"{out}/sf{rank}.pt")
    dist.destroy_process_group()


def test_segformer_route_reduces_every_bucket_in_every_step(tmp_path, monkeypatch, f64_work_dtype):
    port = _free_port()
    mp.spawn(_segformer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "sf0.pt"), torch.load(tmp_path / "sf1.pt")
    assert torch.equal(r0["flat"], r1["flat"])  # ranks that skipped an all-reduce would drift apart from step 2 on
    sys.path.insert(0, str(ROOT / "tests"))
    tr = _segformer_trainer(monkeypatch)
    raw, t = _segformer_data()
    losses = [float(tr.step(raw, t)) for _ in range(3)]
    assert torch.allclose(tr.flat, r0["flat"], atol=1e-9, rtol=1e-6)
    for i in range(3):
        assert abs(losses[i] - 0.5 * (r0["losses"][i] + r1["losses"][i])) < 1e-8


# ---------------------------------------------------------------------------------------------
# tile-sharded sliding-window inference (BASELINE configs[4]): windows dealt round-robin to the ranks, one all-reduce
# ---------------------------------------------------------------------------------------------
def _infer_model():
    import cpu_kernel_emulation as emu
    from gdl_b200.models.unetpp import UnetPlusPlus
    emu.set_work_dtype(torch.float64)
    torch.manual_seed(3)
    m = UnetPlusPlus("resnet18", in_channels=3, classes=4, compute_dtype=torch.float64).double().eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
    return m


def _raster():
    return torch.randint(0, 256, (100, 70, 3), generator=torch.Generator().manual_seed(8), dtype=torch.uint8)


def _infer_worker(rank, world, port, out):
    for p in (ROOT, ROOT / "geo-deep-learning_b200", ROOT / "tests"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)  # two workers on one host: full-width OpenMP teams spin against each other
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_kernel_emulation as emu
    from gdl_b200.inference import SlidingWindowSegmenter
    emu.install_global()
    seg = SlidingWindowSegmenter(_infer_model(), tile=64, stride=32, batch=2, mean=[0.4, 0.5, 0.6], std=[0.2, 0.25, 0.3])
    assert seg.world == world and seg.rank == rank
    logits = seg.logits(_raster())
    torch.save({"logits": logits, "windows": seg.windows_done}, f"{out}/infer{rank}.pt")
    dist.destroy_process_group()


def test_sliding_window_sharded_over_two_ranks(tmp_path, monkeypatch, f64_work_dtype):
    import cpu_kernel_emulation as emu
    from gdl_b200.inference import SlidingWindowSegmenter, window_origins
    port = _free_port()
    mp.spawn(_infer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "infer0.pt"), torch.load(tmp_path / "infer1.pt")
    nwin = len(window_origins(100, 64, 32)) * len(window_origins(70, 64, 32))
    assert r0["windows"] + r1["windows"] == nwin and abs(r0["windows"] - r1["windows"]) <= 1  # dealt round-robin
    assert torch.equal(r0["logits"], r1["logits"])  # every rank holds the full sum after the all-reduce
    sys.path.insert(0, str(ROOT / "tests"))
    emu.install(monkeypatch)
    single = SlidingWindowSegmenter(_infer_model(), tile=64, stride=32, batch=2, mean=[0.4, 0.5, 0.6], std=[0.2, 0.25, 0.3])
    want = single.logits(_raster())
    assert single.windows_done == nwin
    assert torch.allclose(r0["logits"], want, atol=1e-5, rtol=1e-5)  # fp32 accumulator, different summation order
