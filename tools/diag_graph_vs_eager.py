"""Adjudicates tests/test_unetpp_gpu.py::test_cuda_graph_step_equals_eager_step:
is the captured step different arithmetic from the eager one, or is the gap run-to-run reduction noise?

Compares the flat gradient after ONE forward_backward (same weights, same tiles) between
  (a) two eager runs                      -> run-to-run noise of the atomically accumulated sums
  (b) an eager run and a graph replay     -> must be no larger than (a)
and the running statistics after it.  Prints one JSON line.
  python tools/diag_graph_vs_eager.py [--hw 128] [--batch 8] [--enc resnet18]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "geo-deep-learning_b200"):
    sys.path.insert(0, str(p))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--enc", default="resnet18")
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    from gdl_b200 import ops
    from gdl_b200.models.unetpp import UnetPlusPlus
    from gdl_b200.trainer import FusedTrainer

    def fresh():
        torch.manual_seed(5)
        m = UnetPlusPlus(a.enc, in_channels=3, classes=5).cuda().train()
        return m, FusedTrainer(m, ops.LossSpec(1.0, 0.0, ignore_index=-100), lr=2e-3, mean=[0.5] * 3, std=[0.2] * 3)

    g = torch.Generator().manual_seed(6)
    t = torch.randint(0, 5, (a.batch, 4, 4), generator=g).repeat_interleave(a.hw // 4, 1).repeat_interleave(a.hw // 4, 2).cuda()
    raw = (t.unsqueeze(-1) * 50 + torch.randint(0, 30, (a.batch, a.hw, a.hw, 3), generator=g).cuda()).to(torch.uint8)

    def rel(x, y):
        return ((x - y).norm() / (y.norm() + 1e-30)).item()

    out = {"enc": a.enc, "hw": a.hw, "batch": a.batch, "deterministic": ops.deterministic() if hasattr(ops, "deterministic") else None}
    _, tr = fresh()
    l0 = tr.forward_backward(raw, t).item()
    g0 = tr.gflat.clone()
    l1 = tr.forward_backward(raw, t).item()
    g1 = tr.gflat.clone()
    out["eager_vs_eager"] = {"loss": [l0, l1], "grad_rel": rel(g1, g0), "grad_maxabs": (g1 - g0).abs().max().item(),
                             "bit_equal": bool(torch.equal(g0, g1))}

    m2, tr2 = fresh()
    s_raw, s_t = raw.clone(), t.clone()
    tr2.forward_backward(s_raw, s_t)  # warm-up (lazy init); running stats move, gradients do not depend on them in train mode
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        lg = tr2.forward_backward(s_raw, s_t)
    graph.replay()
    torch.cuda.synchronize()
    gg = tr2.gflat.clone()
    graph.replay()
    torch.cuda.synchronize()
    gg2 = tr2.gflat.clone()
    out["graph_vs_eager"] = {"loss": [l0, lg.item()], "grad_rel": rel(gg, g0), "grad_maxabs": (gg - g0).abs().max().item(),
                             "bit_equal": bool(torch.equal(gg, g0))}
    out["graph_vs_graph"] = {"grad_rel": rel(gg2, gg), "bit_equal": bool(torch.equal(gg2, gg))}

    # trajectories: eager, eager again, graph
    traj = {}
    for name, cg in (("eager_a", False), ("eager_b", False), ("graph", True)):
        m, tr = fresh()
        tr.cuda_graph = cg
        traj[name] = [tr.step(raw, t).item() for _ in range(a.steps)]
        traj[name + "_flat_norm"] = tr.flat.norm().item()
    out["trajectories"] = traj
    print(json.dumps(out))


if __name__ == "__main__":
    main()
