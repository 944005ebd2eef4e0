#!/bin/bash
# single GPU: sliding-window inference, window-batch forward replayed from a CUDA graph (--cuda-graph 3) vs eager launches (2);
# then the -m gpu suite once more (final tree)
mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],1), round(d["e2e"]["value"],1), d["config"]["cuda_graph"], d["gpu_launches"], d["clocks"])'
for cg in 3 2; do
  echo "=== segformer_b5_infer raster 6000 cuda-graph=$cg"
  timeout 600 python bench.py --workload segformer_b5_infer --raster 6000 --steps 2 --warmup 1 --cuda-graph $cg --no-cpu-baseline --no-library-baseline 2>gpurun_out/infer.err | tee gpurun_out/ri_bench_infer_cg$cg.json | python -c "$show"
  tail -2 gpurun_out/infer.err
done
echo "=== pytest -m gpu -x"
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/ri_pytest_gpu_full.log 2>&1; tail -3 gpurun_out/ri_pytest_gpu_full.log
