#!/bin/bash
# single GPU: programmatic dependent launch (option "pdl") — the bitwise tests, then A/B of the bench workloads with GDL_PDL=0/1
mkdir -p gpurun_out
echo "=== pytest tests/test_pdl_gpu.py"
SECONDS=0
timeout 300 python -m pytest tests/test_pdl_gpu.py -x -q -p no:cacheprovider > gpurun_out/rp_pytest_pdl.log 2>&1; tail -15 gpurun_out/rp_pytest_pdl.log; echo "wall=${SECONDS}s"
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "pdl", d["config"]["pdl"], "launches", d["gpu_launches"], d["clocks"])'
for wl in segformer_b2 unetpp_r50; do
  for pdl in 0 1; do
    echo "=== $wl GDL_PDL=$pdl"
    GDL_PDL=$pdl timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --workloads headline 2>gpurun_out/pdl.err | tee gpurun_out/rp_bench_${wl}_pdl$pdl.json | python -c "$show"
    tail -2 gpurun_out/pdl.err
  done
done
