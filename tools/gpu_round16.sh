#!/bin/bash
# run 16: sliding-window inference (test + bench on a 4096^2 and the 10000^2 raster)
mkdir -p gpurun_out
echo "=== inference test"
timeout 600 python -m pytest tests/test_zz1_inference_gpu.py -m gpu -q --no-header -rA -p no:cacheprovider 2>&1 | grep -E "passed|failed|error|sliding|class agreement|Error|assert " | head -12
echo "=== bench infer 4096"; timeout 600 python bench.py --workload segformer_b5_infer --raster 4096 --steps 3 --warmup 1 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_infer_4096.json | cut -c1-700
echo "=== bench infer 10000"; timeout 900 python bench.py --workload segformer_b5_infer --steps 2 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_infer_10000.json | cut -c1-300
tail -5 gpurun_out/bench.err
