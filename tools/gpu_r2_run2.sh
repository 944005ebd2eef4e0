#!/bin/bash
# round 2, GPU call 2: deterministic reductions — tests, graph-vs-eager diag, cost A/B (GDL_DETERMINISTIC=0/1)
mkdir -p gpurun_out
echo "=== determinism tests"
timeout 600 python -m pytest tests/test_determinism_gpu.py -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/r2_determinism.log 2>&1
grep -E " passed| failed| error" gpurun_out/r2_determinism.log | tail -1; grep -E "^(FAILED|ERROR)" gpurun_out/r2_determinism.log | head -20
echo "=== diag"
timeout 300 python tools/diag_graph_vs_eager.py --hw 64 --batch 4 2>gpurun_out/diag.err | tee gpurun_out/r2_diag_64.json | cut -c1-900
timeout 300 python tools/diag_graph_vs_eager.py --hw 128 --batch 8 2>>gpurun_out/diag.err | tee gpurun_out/r2_diag_128.json | cut -c1-900
tail -3 gpurun_out/diag.err
echo "=== full pytest -m gpu (as the driver runs it)"
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_full.log
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d.get("roofline",{}).get("wgrad"), d["gpu_launches"])'
for det in 1 0; do
  echo "=== bench unetpp det=$det"; GDL_DETERMINISTIC=$det timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --table gpurun_out/r2_conv_table_unetpp_det$det.json 2>gpurun_out/bench.err | tee gpurun_out/r2_bench_unetpp_det$det.json | python -c "$show"
  echo "=== bench segformer det=$det"; GDL_DETERMINISTIC=$det timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/r2_bench_sf_det$det.json | python -c "$show"
done
echo "=== bench dofa"; timeout 600 python bench.py --workload dofa_base --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/r2_bench_dofa.json | python -c "$show"
tail -5 gpurun_out/bench.err
python __graft_entry__.py smoke 2>&1 | tail -3
