"""torchrun check of the NVLink peer exchange (gdl_p2p_allreduce_sums) against NCCL: values, bit-identity across ranks,
slot reuse over many back-to-back calls of varying length, and replay from a CUDA graph.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_exchange_check.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "geo-deep-learning_b200"))


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from gdl_b200 import ops
    ex = ops.P2PExchange(dist.group.WORLD, dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    worst = 0.0
    for it in range(200):
        n = [128, 512, 4096, 8192, 64, 2 * 768][it % 6]
        v = torch.randn(n, generator=g, device=dev)
        ref = v.clone()
        dist.all_reduce(ref)
        got = ex.all_reduce_(v.clone())
        worst = max(worst, ((got - ref).abs().max() / ref.abs().max()).item())
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        assert all(torch.equal(gathered[0], t) for t in gathered), f"ranks disagree at call {it}"
    assert worst < 1e-6, worst
    # CUDA graph: three exchanges per replay on static buffers
    bufs = [torch.zeros(n, device=dev) for n in (256, 4096, 130)]
    for b in bufs:
        ex.all_reduce_(b)  # warm-up outside capture
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    src = [torch.zeros_like(b) for b in bufs]
    with torch.cuda.graph(graph):
        for b, s_ in zip(bufs, src):
            b.copy_(s_)
            ex.all_reduce_(b)
    for rep in range(20):
        for s_ in src:
            s_.copy_(torch.randn(s_.shape, generator=g, device=dev))
        graph.replay()
        torch.cuda.synchronize()
        for b, s_ in zip(bufs, src):
            ref = s_.clone()
            dist.all_reduce(ref)
            assert ((b - ref).abs().max() / ref.abs().max()).item() < 1e-6, rep
    # latency
    v = torch.randn(4096, device=dev)
    for fn, name in ((lambda: ex.all_reduce_(v), "nvlink_p2p"), (lambda: dist.all_reduce(v), "nccl")):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"{name}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per 4096-float exchange at N={world} (eager launches)")
    if rank == 0:
        print(f"p2p exchange check OK at N={world}: worst relative deviation from NCCL {worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
