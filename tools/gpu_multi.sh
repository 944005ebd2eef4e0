#!/bin/bash
# N-GPU validation of the data-parallel path (NCCL all-reduce of the flat gradient + SyncBN statistics)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@"; }
echo "=== unetpp N=$N (sync_bn)"; run --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/multi.err | tee gpurun_out/bench_n${N}_unetpp.json | cut -c1-420
echo "=== unetpp N=$N (no sync_bn)"; run --steps 6 --warmup 3 --no-cpu-baseline --sync-bn 0 2>>gpurun_out/multi.err | tee gpurun_out/bench_n${N}_unetpp_nosyncbn.json | cut -c1-200
echo "=== segformer N=$N"; run --workload segformer_b2 --steps 6 --warmup 3 --no-cpu-baseline 2>>gpurun_out/multi.err | tee gpurun_out/bench_n${N}_segformer.json | cut -c1-200
echo "=== dofa N=$N"; run --workload dofa_base --steps 6 --warmup 3 --no-cpu-baseline 2>>gpurun_out/multi.err | tee gpurun_out/bench_n${N}_dofa.json | cut -c1-200
echo "=== reference arm under torchrun"; run --impl reference --steps 1 --warmup 0 2>>gpurun_out/multi.err | cut -c1-200
echo "=== N=1 same box"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>>gpurun_out/multi.err | tee gpurun_out/bench_n1_samebox.json | cut -c1-200
tail -8 gpurun_out/multi.err
