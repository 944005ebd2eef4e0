#!/bin/bash
mkdir -p gpurun_out
echo "=== shift probe"; timeout 300 python tools/probe_shift.py > gpurun_out/probe_shift.log 2>&1; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/probe_shift.json'))
    for mode in (0,1):
        for bo in (0,1):
            print('mode',mode,'bo',bo,[ (x['shift'], round(x['relerr'],4)) for x in r if x['mode']==mode and x['bo']==bo])
except Exception as e:
    print('probe failed', e); print(open('gpurun_out/probe_shift.log').read()[-2000:])
PY
echo "=== pytest (segformer + transformer kernels + unetpp)"
timeout 1200 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider -k "segformer or transformer or unetpp" > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -20
grep -E "logits rel err|worst grad|argmax agreement|losses" gpurun_out/pytest_gpu_full.log | head -40
grep -E "Error|error" gpurun_out/pytest_gpu_full.log | sort | uniq -c | sort -rn | head -10
ls -la gpurun_out | head
