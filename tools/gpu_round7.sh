#!/bin/bash
mkdir -p gpurun_out
echo "=== conv epilogue A/B"; timeout 600 python tools/bench_conv_ab.py 2>&1 | tail -14
echo "=== pytest (upernet + conv tests, both epilogues via default)"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -30
grep -E "^\[(96|768|128) " gpurun_out/pytest_gpu_full.log | head
for epi in 1 0; do
echo "=== bench unetpp epilogue=$epi"; GDL_CONV_EPILOGUE=$epi timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_epi$epi.json 2>gpurun_out/bench.err | tee gpurun_out/bench_epi$epi.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['wgrad']['achieved'])"
echo "=== bench segformer epilogue=$epi"; GDL_CONV_EPILOGUE=$epi timeout 600 python bench.py --workload segformer_b2 --steps 6 --warmup 3 --no-cpu-baseline --table gpurun_out/conv_table_sf_epi$epi.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf_epi$epi.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['wgrad']['achieved'])"
done
tail -5 gpurun_out/bench.err
echo "=== ncu launch list (segformer, eager)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1400 --csv --log-file gpurun_out/launches_segformer.csv python bench.py --workload segformer_b2 --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_launch_bench_sf.log 2>&1; tail -1 gpurun_out/ncu_launch_bench_sf.log | cut -c1-200
ls gpurun_out | head -30
