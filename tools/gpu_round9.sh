#!/bin/bash
# run 9: rows kernel (weight-stationary 3x3, Cout<=64), grad_gather/bn_stats MLP rewrite, DOFA epilogue fix
mkdir -p gpurun_out
echo "=== rows kernel tests (own process: a trap here must not poison the rest)"
timeout 600 python -m pytest tests/test_conv_rows_gpu.py -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_rows.log 2>&1; rc=$?
grep -E "passed|failed|error" gpurun_out/pytest_rows.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/pytest_rows.log | head -12
if [ $rc != 0 ]; then echo "ROWS KERNEL FAILED -> GDL_CONV_ROWS=0 for the rest"; export GDL_CONV_ROWS=0; fi
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -30
grep -E "dofa (feature|seg)|dofa_base tap|all gradients" gpurun_out/pytest_gpu_full.log | head -20
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["wgrad"]["achieved"], d["gpu_launches"])'
for rows in ${GDL_CONV_ROWS:-1} 0; do
echo "=== bench unetpp conv_rows=$rows"; GDL_CONV_ROWS=$rows timeout 600 python bench.py --steps 8 --warmup 3 $( [ $rows = 0 ] && echo --no-cpu-baseline ) --table gpurun_out/conv_table_rows$rows.json 2>gpurun_out/bench.err | tee gpurun_out/bench_rows$rows.json | python -c "$show"
done
echo "=== bench segformer"; timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --table gpurun_out/conv_table_sf.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf.json | python -c "$show"
echo "=== bench dofa"; timeout 900 python bench.py --workload dofa_base --steps 6 --warmup 3 --table gpurun_out/conv_table_dofa.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa.json | python -c "$show"
tail -5 gpurun_out/bench.err
echo "=== ncu dram+time per launch (unetpp, eager)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1700 -c 900 --csv --log-file gpurun_out/launches_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_launch_bench.log 2>&1; tail -1 gpurun_out/ncu_launch_bench.log | cut -c1-200
echo "=== ncu full: conv3x3_rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_rows_kernel -s 10 -c 3 -o gpurun_out/prof_r9_conv_rows -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_full_rows.log 2>&1; tail -1 gpurun_out/ncu_full_rows.log | cut -c1-120
ls -la gpurun_out | tail -12
