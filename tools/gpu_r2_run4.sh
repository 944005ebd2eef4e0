#!/bin/bash
# round 2, GPU call 4: fused head, BASELINE-shape parity, smoke, the full default bench line (sub-workloads + baselines)
mkdir -p gpurun_out
for f in tests/test_upsample_head_gpu.py tests/test_baseline_shapes_gpu.py tests/test_unetpp_gpu.py tests/test_segformer_gpu.py tests/test_dofa_gpu.py tests/test_zz1_inference_gpu.py; do
  b=$(basename "$f" .py)
  timeout 1200 python -m pytest "$f" -m gpu -q --no-header -rA -p no:cacheprovider > "gpurun_out/r4_$b.log" 2>&1
  echo "$b: $(grep -E ' passed| failed| error' "gpurun_out/r4_$b.log" | tail -1)"
  grep -E "^(FAILED|ERROR)|^E  " "gpurun_out/r4_$b.log" | head -12
done
grep -hE "rel err|worst grad|argmax agreement|grad rel err|rel diff|B=32" gpurun_out/r4_test_baseline_shapes_gpu.log gpurun_out/r4_test_upsample_head_gpu.log | head -40
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -4
echo "=== full default bench (as the driver runs it)"
/usr/bin/time -v python bench.py > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err; echo "rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r4_bench_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4_bench_default.json').read().strip().splitlines()[-1])
def show(k,l):
    print(k, 'value', round(l['value'],1), 'ms', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value'],1), 'roof', round(l['roofline']['frac'],3), 'lib', l.get('library_baseline'), 'clk', l['clocks'])
show('headline', d)
print('wgrad', d['roofline']['wgrad'], 'whole', d['roofline'].get('whole_step'), 'cpu', d['cpu_baseline'])
for k,l in d.get('workloads',{}).items(): show(k,l)
P
tail -5 gpurun_out/r4_bench_default.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
