#!/bin/bash
# round 2, 8-GPU call: the driver's scaling command at N=8 (default: whole step incl. NCCL collectives in one CUDA graph)
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@"; }
echo "=== default line at N=$N"; SECONDS=0
run --steps 10 --warmup 3 2>gpurun_out/n8.err > gpurun_out/rn_bench_n${N}_default.json; echo "rc=$? wall=${SECONDS}s"
python - <<P
import json
d=json.loads(open('gpurun_out/rn_bench_n${N}_default.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['config']['cuda_graph'], d['clocks'])
for k,l in d.get('workloads',{}).items(): print(k, round(l['value'],1), round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value'],1), l['clocks'])
P
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/n8.err | tail -5
echo "=== headline, eager launches at N=$N"
run --steps 10 --warmup 3 --cuda-graph 1 --workloads headline --no-cpu-baseline 2>gpurun_out/n8.err | tee gpurun_out/rn_bench_n${N}_unetpp_eager.json | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), d["config"]["cuda_graph"])'
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/n8.err | tail -3
