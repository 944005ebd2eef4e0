#!/bin/bash
# single GPU: the whole -m gpu suite (every failure listed, no -x) and smoke()
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/rs_pytest_gpu_full.log 2>&1
tail -12 gpurun_out/rs_pytest_gpu_full.log; echo "wall=${SECONDS}s"
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
