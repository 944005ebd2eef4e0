#!/bin/bash
# round 2, GPU call 13 (final single-GPU pass): the driver's three commands, the HBM microbench, and the ncu launch list of the bench command
mkdir -p gpurun_out
echo "=== pytest -m gpu -x (as the driver runs it)"
SECONDS=0
timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r13_pytest_gpu_full.log 2>&1
tail -4 gpurun_out/r13_pytest_gpu_full.log; echo "wall=${SECONDS}s"
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== HBM-bound kernel microbench"
timeout 900 python tools/bench_hbm_kernels.py --out gpurun_out/r13_hbm_kernels.json 2>gpurun_out/hbm.err | grep -E "bn_|grad_gather|normalize_to" | cut -c1-220
tail -2 gpurun_out/hbm.err
echo "=== full default bench (as the driver runs it)"
SECONDS=0
python bench.py > gpurun_out/r13_bench_default.json 2> gpurun_out/r13_bench_default.err; echo "rc=$? wall=${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r13_bench_default.json').read().strip().splitlines()[-1])
def show(k,l):
    print(k, 'value', round(l['value'],1), 'ms', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value'],1), 'roof', round(l['roofline']['frac'],3), 'lib', round((l.get('library_baseline') or {}).get('value',0),1), 'launches', l['gpu_launches'], 'clk', l['clocks'])
show('headline', d)
print('wgrad', d['roofline']['wgrad'], 'whole', d['roofline'].get('whole_step'), 'traffic', d['roofline']['traffic'], d['roofline']['traffic_source'], 'cpu', (d['cpu_baseline'] or {}).get('value'))
for k,l in d.get('workloads',{}).items(): show(k,l)
P
tail -3 gpurun_out/r13_bench_default.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
echo "=== ncu: time + DRAM bytes of every launch of ONE eager UNet++ step at the bench batch (B=32), final code"
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 9700 -c 5200 --csv \
  --log-file gpurun_out/r13_ncu_unetpp_b32_launches.csv python bench.py --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline --no-library-baseline --workloads headline > gpurun_out/r13_ncu_bench.log 2>&1
wc -l gpurun_out/r13_ncu_unetpp_b32_launches.csv
