#!/bin/bash
# round 2, GPU call 9: fused BN statistics (tests + A/B), bf16x2 accuracy mode, then the whole -m gpu suite as the driver runs it
mkdir -p gpurun_out
for f in tests/test_conv_bn_fused_gpu.py tests/test_fp32_split_gpu.py; do
  b=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --no-header -rA -p no:cacheprovider > "gpurun_out/r9_$b.log" 2>&1
  echo "$b: $(grep -E ' passed| failed| error' "gpurun_out/r9_$b.log" | tail -1)"
  grep -E "^(FAILED|ERROR)|^E  " "gpurun_out/r9_$b.log" | head -12
done
grep -h "bf16x2 fwd\|max |fused" gpurun_out/r9_*.log | head -20
echo "=== full pytest -m gpu (as the driver runs it)"
SECONDS=0
timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r9_pytest_gpu_full.log 2>&1
tail -4 gpurun_out/r9_pytest_gpu_full.log; echo "wall=${SECONDS}s"
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), round(d["roofline"]["wgrad"]["achieved"],1), d["gpu_launches"])'
for bn in 1 0; do
  echo "=== unetpp bn_fused=$bn"; GDL_BN_FUSED=$bn timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --workloads headline --table gpurun_out/r9_conv_table_unetpp_bn$bn.json 2>gpurun_out/bench.err | tee gpurun_out/r9_bench_unetpp_bn$bn.json | python -c "$show"
  echo "=== dofa bn_fused=$bn"; GDL_BN_FUSED=$bn timeout 600 python bench.py --workload dofa_base --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline 2>>gpurun_out/bench.err | tee gpurun_out/r9_bench_dofa_bn$bn.json | python -c "$show"
done
tail -3 gpurun_out/bench.err
python __graft_entry__.py smoke 2>&1 | tail -2
