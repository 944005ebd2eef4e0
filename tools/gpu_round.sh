#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 600 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json; tail -20 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1400} -c ${NCU_COUNT:-1500} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; tail -3 gpurun_out/ncu_launch_bench.log
echo "=== ncu full (conv fwd / wgrad)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 60 -c 3 -o gpurun_out/prof_conv_fwd -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_fwd.log 2>&1; tail -2 gpurun_out/ncu_full_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 30 -c 3 -o gpurun_out/prof_conv_wgrad -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_wgrad.log 2>&1; tail -2 gpurun_out/ncu_full_wgrad.log
fi
ls -la gpurun_out | head -30
