#!/bin/bash
# round 2, GPU call 14: lean grad_gather variants (tests, microbench, bench), then the ncu launch list of the bench command
mkdir -p gpurun_out
for f in tests/test_determinism_gpu.py tests/test_kernels_gpu.py tests/test_unetpp_gpu.py; do
  b=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --no-header -p no:cacheprovider > "gpurun_out/r14_$b.log" 2>&1
  echo "$b: $(grep -E ' passed| failed| error' "gpurun_out/r14_$b.log" | tail -1)"; grep -E "^(FAILED|ERROR)|^E  " "gpurun_out/r14_$b.log" | head -8
done
echo "=== HBM-bound kernel microbench"
timeout 900 python tools/bench_hbm_kernels.py --out gpurun_out/r14_hbm_kernels.json 2>gpurun_out/hbm.err | grep -E "grad_gather|bn_stats" | cut -c1-200
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), round(d["roofline"]["wgrad"]["achieved"],1), d["gpu_launches"])'
echo "=== bench unetpp"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --workloads headline 2>gpurun_out/bench.err | tee gpurun_out/r14_bench_unetpp.json | python -c "$show"
echo "=== bench dofa"; timeout 900 python bench.py --workload dofa_base --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline 2>>gpurun_out/bench.err | tee gpurun_out/r14_bench_dofa.json | python -c "$show"
tail -2 gpurun_out/bench.err
echo "=== ncu: time + DRAM bytes of every launch of ~one eager UNet++ step at the bench batch (B=32), final code"
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 1500 -c 1300 --csv \
  --log-file gpurun_out/r14_ncu_unetpp_b32_launches.csv python bench.py --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline --no-library-baseline --workloads headline > gpurun_out/r14_ncu_bench.log 2>&1
wc -l gpurun_out/r14_ncu_unetpp_b32_launches.csv
