#!/bin/bash
# N GPUs: SegFormer-B2 training (BASELINE configs[2]) with the default options — captured step incl. the bucketed gradient
# all-reduce of the model.backward route, SyncBN exchange, folded decoder
N=${1:-2}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload segformer_b2 --steps 10 --warmup 3 --workloads headline --no-cpu-baseline --no-library-baseline 2>gpurun_out/nsf.err | tee gpurun_out/rg_bench_n${N}_sf_b2.json | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["config"]["syncbn_exchange"], d["config"]["cuda_graph"], d["config"].get("decoder_folded"), d["clocks"])'
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/nsf.err | tail -4
