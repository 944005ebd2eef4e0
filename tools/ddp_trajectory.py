"""torchrun: 6 fused UNet++-R18 steps with SyncBN under a CUDA graph; prints rank 0's losses and a checksum of the parameters
(used to compare GDL_OVERLAP_ALLREDUCE=0/1 and GDL_P2P_SYNCBN=0/1: the trajectories must agree)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "geo-deep-learning_b200"))


def main() -> None:
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from gdl_b200 import ops
    from gdl_b200.models.unetpp import UnetPlusPlus
    from gdl_b200.trainer import FusedTrainer
    torch.manual_seed(0)
    m = UnetPlusPlus("resnet18", in_channels=3, classes=5).to(dev).train()
    tr = FusedTrainer(m, ops.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-3, mean=[0.5] * 3, std=[0.2] * 3, sync_bn=True, cuda_graph=True)
    g = torch.Generator().manual_seed(10 + rank)
    t = torch.randint(0, 5, (4, 4, 4), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2).to(dev)
    raw = (t.unsqueeze(-1) * 50 + torch.randint(0, 30, (4, 128, 128, 3), generator=g).to(dev)).to(torch.uint8)
    losses = [round(tr.step(raw, t).item(), 6) for _ in range(6)]
    chk = tr.flat.double().sum().item()
    allchk = [None] * dist.get_world_size()
    dist.all_gather_object(allchk, chk)
    if rank == 0:
        print(f"overlap={ops.option('overlap_allreduce')} exchange={tr.bn_exchange_kind} losses {losses} param checksum {chk:.9f} "
              f"ranks identical: {len(set(allchk)) == 1}")
    # release the symmetric-memory exchange (and everything else that lives in the process group) BEFORE the group goes away:
    # torn down at interpreter exit, after destroy_process_group, it blocked until the launcher's timeout (round 2, run 16)
    del tr, m
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
