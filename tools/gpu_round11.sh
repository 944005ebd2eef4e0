#!/bin/bash
# run 11: paired-tap row-streaming wgrad kernel (Cout 16/32/64)
mkdir -p gpurun_out
echo "=== wgrad rows tests (own process)"
timeout 600 python -m pytest tests/test_wgrad_rows_gpu.py -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1; rc=$?
grep -E "passed|failed|error" gpurun_out/pytest_new.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/pytest_new.log | head -20
if [ $rc != 0 ]; then echo "WGRAD ROWS FAILED -> GDL_WGRAD_ROWS=0 for the rest"; export GDL_WGRAD_ROWS=0; fi
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -30
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["wgrad"]["achieved"], d["gpu_launches"])'
for wr in ${GDL_WGRAD_ROWS:-1} 0; do
echo "=== bench unetpp epilogue=2 wgrad_rows=$wr"; GDL_CONV_EPILOGUE=2 GDL_WGRAD_ROWS=$wr timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_wr$wr.json 2>gpurun_out/bench.err | tee gpurun_out/bench_wr$wr.json | python -c "$show"
done
tail -5 gpurun_out/bench.err
echo "=== ncu full: wgrad3x3_rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_rows_kernel -s 8 -c 3 -o gpurun_out/prof_r11_wgrad_rows -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_full_wr.log 2>&1; tail -1 gpurun_out/ncu_full_wr.log | cut -c1-120
ls -la gpurun_out | tail -6
