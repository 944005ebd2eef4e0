#!/bin/bash
# round 2, GPU call 5: new wgrad split rule + in-place weight refresh: tests, A/B, the full default bench line
mkdir -p gpurun_out
for f in tests/test_determinism_gpu.py tests/test_baseline_shapes_gpu.py tests/test_unetpp_gpu.py tests/test_kernels_gpu.py tests/test_wgrad_rows_gpu.py tests/test_grouped_gemm_gpu.py tests/test_segformer_gpu.py; do
  b=$(basename "$f" .py)
  timeout 1200 python -m pytest "$f" -m gpu -q --no-header -rA -p no:cacheprovider > "gpurun_out/r5_$b.log" 2>&1
  echo "$b: $(grep -E ' passed| failed| error' "gpurun_out/r5_$b.log" | tail -1)"
  grep -E "^(FAILED|ERROR)|^E  " "gpurun_out/r5_$b.log" | head -12
done
grep -h "repack_in_place\|B=32" gpurun_out/r5_*.log | head
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), d["roofline"]["wgrad"]["achieved"], d["roofline"]["wgrad"]["share_of_step"], d["gpu_launches"])'
for sched in 1 0; do
  echo "=== unetpp wgrad_sched=$sched"; GDL_WGRAD_SCHED=$sched timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --workloads headline --table gpurun_out/r5_conv_table_unetpp_sched$sched.json 2>gpurun_out/bench.err | tee gpurun_out/r5_bench_unetpp_sched$sched.json | python -c "$show"
  echo "=== segformer wgrad_sched=$sched"; GDL_WGRAD_SCHED=$sched timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline 2>>gpurun_out/bench.err | tee gpurun_out/r5_bench_sf_sched$sched.json | python -c "$show"
done
echo "=== segformer det=0"; GDL_DETERMINISTIC=0 timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline 2>>gpurun_out/bench.err | tee gpurun_out/r5_bench_sf_det0.json | python -c "$show"
echo "=== segformer fused_head=0"; GDL_FUSED_HEAD=0 timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline 2>>gpurun_out/bench.err | tee gpurun_out/r5_bench_sf_head0.json | python -c "$show"
tail -3 gpurun_out/bench.err
echo "=== full default bench (as the driver runs it)"
SECONDS=0
python bench.py > gpurun_out/r5_bench_default.json 2> gpurun_out/r5_bench_default.err; echo "rc=$? wall=${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r5_bench_default.json').read().strip().splitlines()[-1])
def show(k,l):
    print(k, 'value', round(l['value'],1), 'ms', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value'],1), 'roof', round(l['roofline']['frac'],3), 'lib', (l.get('library_baseline') or {}).get('value'), (l.get('library_baseline') or {}).get('unavailable'), 'clk', l['clocks'])
show('headline', d)
print('wgrad', d['roofline']['wgrad'], 'whole', d['roofline'].get('whole_step'), 'cpu', (d['cpu_baseline'] or {}).get('value'))
for k,l in d.get('workloads',{}).items(): show(k,l)
P
tail -5 gpurun_out/r5_bench_default.err
