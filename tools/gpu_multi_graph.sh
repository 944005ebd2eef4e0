#!/bin/bash
# N-GPU: whole-step CUDA graph with the NCCL collectives captured (--cuda-graph 2) vs eager
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@"; }
echo "=== unetpp N=$N graph"; run --steps 6 --warmup 3 --no-cpu-baseline --cuda-graph 2 2>gpurun_out/multi_graph.err | tee gpurun_out/bench_n${N}_unetpp_graph.json | cut -c1-300; echo "rc=$?"
echo "=== segformer N=$N graph"; run --workload segformer_b2 --steps 6 --warmup 3 --no-cpu-baseline --cuda-graph 2 2>>gpurun_out/multi_graph.err | tee gpurun_out/bench_n${N}_segformer_graph.json | cut -c1-300
echo "=== dofa N=$N graph"; run --workload dofa_base --steps 6 --warmup 3 --no-cpu-baseline --cuda-graph 2 2>>gpurun_out/multi_graph.err | tee gpurun_out/bench_n${N}_dofa_graph.json | cut -c1-300
echo "=== segformer N=$N eager"; run --workload segformer_b2 --steps 6 --warmup 3 --no-cpu-baseline 2>>gpurun_out/multi_graph.err | tee gpurun_out/bench_n${N}_segformer_eager.json | cut -c1-300
echo "=== reference arm under torchrun"; run --impl reference --steps 1 --warmup 0 2>>gpurun_out/multi_graph.err | cut -c1-300
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/multi_graph.err | tail -15
