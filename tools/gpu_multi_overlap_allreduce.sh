#!/bin/bash
# round 2, N-GPU call: bucketed all-reduce overlapped with the backward (graph-captured) vs one all-reduce after it
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --no-cpu-baseline --no-library-baseline --workloads headline "$@"; }
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["config"].get("syncbn_exchange"), d["config"]["cuda_graph"], d["clocks"])'
for ov in 1 0; do
  echo "=== unetpp N=$N overlap_allreduce=$ov (graph)"; GDL_OVERLAP_ALLREDUCE=$ov run --steps 10 --warmup 3 2>gpurun_out/ov.err | tee gpurun_out/ro_bench_n${N}_unetpp_ov$ov.json | python -c "$show"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/ov.err | tail -3
done
echo "=== unetpp N=$N overlap_allreduce=1 (eager)"; GDL_OVERLAP_ALLREDUCE=1 run --steps 10 --warmup 3 --cuda-graph 1 2>gpurun_out/ov.err | python -c "$show"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/ov.err | tail -3
echo "=== trajectory check: 6 steps, overlap on/off must give the same loss sequence"
for ov in 1 0; do
GDL_OVERLAP_ALLREDUCE=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29477 tools/ddp_trajectory.py 2>gpurun_out/ov.err | tail -1
done
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/ov.err | tail -3
