#!/bin/bash
# run 10: TMA-store epilogues (rows kernel + conv_fwd_kernel mode 2), contiguous-front grad_gather / bn_stats
mkdir -p gpurun_out
echo "=== new kernels' tests (own process)"
timeout 600 python -m pytest tests/test_conv_rows_gpu.py tests/test_conv_epilogue_modes_gpu.py -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1; rc=$?
grep -E "passed|failed|error" gpurun_out/pytest_new.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/pytest_new.log | head -20
echo "=== pytest -m gpu (default options)"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -30
echo "=== pytest -m gpu with GDL_CONV_EPILOGUE=2 (model-level parity on the TMA-store epilogue)"
GDL_CONV_EPILOGUE=2 timeout 1500 python -m pytest tests/test_unetpp_gpu.py tests/test_segformer_gpu.py tests/test_dofa_gpu.py tests/test_upernet_gpu.py tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["wgrad"]["achieved"], d["gpu_launches"])'
for cfg in "0 0" "0 1" "2 1"; do set -- $cfg
echo "=== bench unetpp epilogue=$1 rows_tma_store=$2"; GDL_CONV_EPILOGUE=$1 GDL_ROWS_TMA_STORE=$2 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_e$1_r$2.json 2>gpurun_out/bench.err | tee gpurun_out/bench_e$1_r$2.json | python -c "$show"
done
for epi in 0 2; do
echo "=== bench segformer epilogue=$epi"; GDL_CONV_EPILOGUE=$epi timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_sf_e$epi.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf_e$epi.json | python -c "$show"
echo "=== bench dofa epilogue=$epi"; GDL_CONV_EPILOGUE=$epi timeout 900 python bench.py --workload dofa_base --steps 6 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_dofa_e$epi.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa_e$epi.json | python -c "$show"
done
tail -5 gpurun_out/bench.err
echo "=== ncu launch list (unetpp, eager, epilogue=2)"
GDL_CONV_EPILOGUE=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 900 --csv --log-file gpurun_out/launches_e2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_launch_bench.log 2>&1; tail -1 gpurun_out/ncu_launch_bench.log | cut -c1-160
ls -la gpurun_out | tail -8
