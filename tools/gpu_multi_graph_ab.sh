#!/bin/bash
# round 2, N-GPU call: whole-step CUDA graph with the NCCL collectives captured (--cuda-graph 2) vs eager launches (1)
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --no-cpu-baseline --no-library-baseline "$@"; }
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["config"]["cuda_graph"], d["gpu_launches"], d["clocks"])'
for cg in 2 1; do
  echo "=== unetpp N=$N cuda-graph=$cg"; run --steps 8 --warmup 3 --workloads headline --cuda-graph $cg 2>gpurun_out/multi.err | tee gpurun_out/rm_bench_n${N}_unetpp_cg$cg.json | python -c "$show"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/multi.err | tail -4
  echo "=== segformer N=$N cuda-graph=$cg"; run --workload segformer_b2 --steps 8 --warmup 3 --cuda-graph $cg 2>gpurun_out/multi.err | tee gpurun_out/rm_bench_n${N}_sf_cg$cg.json | python -c "$show"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/multi.err | tail -4
done
echo "=== dofa N=$N cuda-graph=2"; run --workload dofa_base --steps 8 --warmup 3 --cuda-graph 2 2>gpurun_out/multi.err | tee gpurun_out/rm_bench_n${N}_dofa_cg2.json | python -c "$show"
echo "=== infer N=$N"; run --workload segformer_b5_infer --raster 6000 --steps 2 --warmup 1 2>gpurun_out/multi.err | tee gpurun_out/rm_bench_n${N}_infer.json | python -c "$show"
echo "=== full default line at N=$N (sub-workloads ride along)"; SECONDS=0; run --steps 6 --warmup 3 --cuda-graph 2 2>gpurun_out/multi.err > gpurun_out/rm_bench_n${N}_default.json; echo "rc=$? wall=${SECONDS}s"; python - <<P
import json
d=json.loads(open('gpurun_out/rm_bench_n${N}_default.json').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), round(d['ms_per_step'],2))
for k,l in d.get('workloads',{}).items(): print(k, round(l['value'],1), round(l['ms_per_step'],2), l['clocks'])
P
echo "=== reference arm under torchrun"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29433 bench.py --gpus $N --impl reference --steps 1 --warmup 0 2>>gpurun_out/multi.err | cut -c1-200
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/multi.err | tail -6
