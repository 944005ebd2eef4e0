#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -30
grep -E "^eager|^graph" gpurun_out/pytest_gpu_full.log | head
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench unetpp (graph)"; timeout 600 python bench.py --steps 8 --warmup 3 --table gpurun_out/conv_table.json 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-1800; tail -5 gpurun_out/bench.err
echo "=== bench unetpp (no graph)"; timeout 600 python bench.py --steps 5 --warmup 3 --cuda-graph 0 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_nograph.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "=== bench segformer_b2 (graph)"; timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --table gpurun_out/conv_table_segformer.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_segformer.json | cut -c1-1800; tail -5 gpurun_out/bench.err
echo "=== bench segformer_b2 (no graph)"; timeout 600 python bench.py --workload segformer_b2 --steps 5 --warmup 3 --cuda-graph 0 --no-cpu-baseline 2>>gpurun_out/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "=== ncu launch list (unetpp, eager)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_launch_bench.log 2>&1; tail -1 gpurun_out/ncu_launch_bench.log | cut -c1-200
echo "=== ncu launch list (segformer, eager)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1400 --csv --log-file gpurun_out/launches_segformer.csv python bench.py --workload segformer_b2 --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_launch_bench_sf.log 2>&1; tail -1 gpurun_out/ncu_launch_bench_sf.log | cut -c1-200
ls -la gpurun_out | head -30
