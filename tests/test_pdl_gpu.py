"""GPU: programmatic dependent launch (option "pdl", csrc/common.cuh) changes WHEN a kernel becomes resident, never what it
computes: every kernel blocks in griddepcontrol.wait until its predecessor has completed.  With the ordered reductions a
training trajectory is bit-reproducible, so the check is exact: eager launches without the attribute == eager launches with
it == the CUDA-graph replay with it (where the dependencies become programmatic graph edges) — losses and every parameter,
for the convolutional (UNet++) and the transformer (SegFormer, fused attention, gradient clipping) step, and for the
graph-replayed sliding-window forward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def pdl_switch():
    from gdl_b200 import ops
    old = ops.option("pdl")
    yield ops
    ops.set_option("pdl", old)


def _trajectories(make_trainer, raw, t, steps, ops):
    losses, flats = {}, {}
    for mode, pdl, graph in (("plain", 0, False), ("pdl_eager", 1, False), ("pdl_graph", 1, True)):
        ops.set_option("pdl", pdl)
        assert ops.option("pdl") == pdl
        tr = make_trainer(graph)
        losses[mode] = [tr.step(raw[i % len(raw)], t).item() for i in range(steps)]
        flats[mode] = tr.flat.clone()
        if graph:
            assert tr._graph is not None and tr.launches_per_step > 100
        torch.cuda.synchronize()
    return losses, flats


def test_unetpp_step_is_bitwise_unchanged_by_pdl(cuda, pdl_switch):
    from test_unetpp_gpu import _models
    from gdl_b200.ops import LossSpec
    from gdl_b200.trainer import FusedTrainer
    ops = pdl_switch
    assert ops.deterministic()
    g = torch.Generator().manual_seed(6)
    t = torch.randint(0, 5, (4, 2, 2), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).cuda()
    raw = [(t.unsqueeze(-1) * 50 + torch.randint(0, 30, (4, 64, 64, 3), generator=g).cuda()).to(torch.uint8) for _ in range(3)]

    def make(graph):
        _, prod = _models("resnet18", 3, 5, seed=5)
        return FusedTrainer(prod.train(), LossSpec(1.0, 0.0, ignore_index=-100), lr=2e-3, mean=[0.5] * 3, std=[0.2] * 3,
                            cuda_graph=graph)
    losses, flats = _trajectories(make, raw, t, 8, ops)
    print("plain", [round(v, 4) for v in losses["plain"]])
    assert losses["plain"][-1] < 0.8 * losses["plain"][0]
    for mode in ("pdl_eager", "pdl_graph"):
        assert losses[mode] == losses["plain"], mode
        assert torch.equal(flats[mode], flats["plain"]), mode


def test_segformer_step_is_bitwise_unchanged_by_pdl(cuda, pdl_switch):
    from test_segformer_gpu import _setup
    from gdl_b200.ops import LossSpec
    from gdl_b200.trainer import FusedTrainer
    ops = pdl_switch
    g = torch.Generator().manual_seed(6)
    t = torch.randint(0, 5, (4, 4, 4), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).cuda()
    raw = [(t.unsqueeze(-1) * 50 + torch.randint(0, 30, (4, 128, 128, 3), generator=g).cuda()).to(torch.uint8)]

    def make(graph):
        prod = _setup("mit_b1", 3, 5, seed=3).train()
        return FusedTrainer(prod, LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.2] * 3,
                            clip_grad_norm=1.0, cuda_graph=graph)
    losses, flats = _trajectories(make, raw, t, 5, ops)
    for mode in ("pdl_eager", "pdl_graph"):
        assert losses[mode] == losses["plain"], mode
        assert torch.equal(flats[mode], flats["plain"]), mode


def test_sliding_window_forward_is_bitwise_unchanged_by_pdl(cuda, pdl_switch):
    from test_segformer_gpu import _setup
    from gdl_b200.inference import SlidingWindowSegmenter
    ops = pdl_switch
    model = _setup("mit_b0", 3, 5, seed=4).eval()
    raster = torch.randint(0, 256, (200, 232, 3), generator=torch.Generator().manual_seed(2), dtype=torch.uint8).cuda()
    out = {}
    for mode, pdl, graph in (("plain", 0, False), ("pdl_eager", 1, False), ("pdl_graph", 1, True)):
        ops.set_option("pdl", pdl)
        seg = SlidingWindowSegmenter(model, tile=64, stride=32, batch=4, mean=[0.4, 0.5, 0.6], std=[0.2, 0.25, 0.3],
                                     cuda_graph=graph)
        out[mode] = seg.logits(raster).clone()
        if graph:
            assert seg._graph is not None
    assert torch.equal(out["pdl_eager"], out["plain"]) and torch.equal(out["pdl_graph"], out["plain"])
