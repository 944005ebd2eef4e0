#!/bin/bash
# round 2, N-GPU call: NVLink peer exchange of the SyncBN statistics — check against NCCL, then the bench A/B
N=${1:-2}
mkdir -p gpurun_out
echo "=== p2p exchange check"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/p2p_exchange_check.py 2>gpurun_out/p2p.err | tail -5
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/p2p.err | tail -8
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --no-cpu-baseline --no-library-baseline --workloads headline "$@"; }
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["config"].get("syncbn_exchange"), d["config"]["cuda_graph"], d["clocks"])'
for p2p in 1 0; do
  echo "=== unetpp N=$N p2p_syncbn=$p2p"; GDL_P2P_SYNCBN=$p2p run --steps 10 --warmup 3 2>gpurun_out/p2p.err | tee gpurun_out/rp_bench_n${N}_unetpp_p2p$p2p.json | python -c "$show"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/p2p.err | tail -3
done
echo "=== dofa N=$N p2p_syncbn=1"; GDL_P2P_SYNCBN=1 run --workload dofa_base --steps 10 --warmup 3 2>gpurun_out/p2p.err | tee gpurun_out/rp_bench_n${N}_dofa_p2p1.json | python -c "$show"
