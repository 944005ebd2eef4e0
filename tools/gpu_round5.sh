#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -20
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench unetpp (halo on)"; timeout 600 python bench.py --steps 8 --warmup 3 --table gpurun_out/conv_table.json 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-1700; tail -5 gpurun_out/bench.err
echo "=== bench unetpp (halo off)"; GDL_CONV_HALO=0 GDL_WGRAD_HALO=0 timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_nohalo.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['wgrad'])"
echo "=== bench segformer_b2"; timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --table gpurun_out/conv_table_segformer.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_segformer.json | cut -c1-1700; tail -5 gpurun_out/bench.err
echo "=== ncu launch list (unetpp)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; tail -1 gpurun_out/ncu_launch_bench.log | cut -c1-200
echo "=== ncu launch list (segformer)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/launches_segformer.csv python bench.py --workload segformer_b2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench_sf.log 2>&1; tail -1 gpurun_out/ncu_launch_bench_sf.log | cut -c1-200
ls -la gpurun_out | head -30
