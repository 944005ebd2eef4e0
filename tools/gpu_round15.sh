#!/bin/bash
# run 15: grouped (all-heads) attention GEMM launches, 128-bit softmax
mkdir -p gpurun_out
echo "=== new tests (own process)"
timeout 900 python -m pytest tests/test_grouped_gemm_gpu.py tests/test_transformer_kernels_gpu.py tests/test_segformer_gpu.py tests/test_dofa_gpu.py -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1; rc=$?
grep -E "passed|failed|error" gpurun_out/pytest_new.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/pytest_new.log | head -20
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -30
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["wgrad"]["achieved"], d["gpu_launches"])'
echo "=== bench segformer"; timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_sf.json 2>gpurun_out/bench.err | tee gpurun_out/bench_sf.json | python -c "$show"
echo "=== bench dofa"; timeout 900 python bench.py --workload dofa_base --steps 6 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_dofa.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa.json | python -c "$show"
tail -5 gpurun_out/bench.err
