"""GPU (>= 2 devices): the NVLink peer exchange of the SyncBN statistics (gdl_p2p_allreduce_sums, ops.P2PExchange) against
NCCL — values, bit-identity across ranks, slot reuse over 200 back-to-back calls of six lengths, replay from a CUDA graph.
Runs tools/p2p_exchange_check.py under torchrun on two devices; skipped on a single-GPU box (the driver's test box),
where the check was run by hand instead (profiles/r02_run12_*: deviation from NCCL 0.0 at N = 2; N = 8 in the bench lines)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_p2p_exchange_matches_nccl_on_two_gpus(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29641", str(ROOT / "tools" / "p2p_exchange_check.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "p2p exchange check OK at N=2" in r.stdout, r.stdout[-2000:]
