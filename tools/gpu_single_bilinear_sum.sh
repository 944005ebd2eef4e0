#!/bin/bash
# single GPU: the new decoder kernel alone — event timing against the copy peak, then one ncu --set full capture
mkdir -p gpurun_out
timeout 60 python tools/bench_bilinear_sum.py --out gpurun_out/rb_bilinear_sum.json 2>&1 | tail -3
timeout 70 ncu --set full --clock-control none --import-source on -k regex:bilinear_sum -s 2 -c 1 -f -o gpurun_out/rb_bilinear_sum python tools/bench_bilinear_sum.py --once > gpurun_out/rb_ncu.log 2>&1; tail -2 gpurun_out/rb_ncu.log
