#!/bin/bash
# One GPU session: parity tests, smoke, bench (+per-launch conv table), optional variants, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -20
grep -E "rel err|worst|mismatch|argmax agreement|product .* autocast|^e[1-5] " gpurun_out/pytest_gpu_full.log | head -60
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 600 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 --table gpurun_out/conv_table.json 2>gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for v in ${WGRAD_L2_VARIANTS:-}; do
  echo "=== bench GDL_WGRAD_L2_MB=$v"
  GDL_WGRAD_L2_MB=$v timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_l2_$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['wgrad'])"
done
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1700} -c ${NCU_COUNT:-900} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-300
fi
if [ "${NCU_FULL:-0}" == "1" ]; then
echo "=== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-conv_fwd_kernel} -s ${NCU_FULL_SKIP:-60} -c 3 -o gpurun_out/prof_${NCU_KERNEL:-conv_fwd_kernel} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
fi
ls -la gpurun_out | head -30
