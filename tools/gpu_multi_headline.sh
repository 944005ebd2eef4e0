#!/bin/bash
# N GPUs: the headline workload only (as the driver's scaling run launches it), default options
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --workloads headline --no-cpu-baseline --no-library-baseline 2>gpurun_out/n8h.err | tee gpurun_out/rh_bench_n${N}_unetpp.json | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["config"]["syncbn_exchange"], d["config"]["cuda_graph"], d["clocks"])'
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/n8h.err | tail -4
