#!/bin/bash
# run 8: DOFA + pixel-packed narrow convs + grad_gather fast path: parity, benches (A/B), dram-traffic pass, ncu full
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -30
grep -E "dofa|attention, 1297" gpurun_out/pytest_gpu_full.log | grep -v PASSED | head -20
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["wgrad"]["achieved"], d["gpu_launches"])'
for pp in 1 0; do
echo "=== bench unetpp pixel_pack=$pp"; GDL_PIXEL_PACK=$pp timeout 600 python bench.py --steps 8 --warmup 3 $( [ $pp = 0 ] && echo --no-cpu-baseline ) --table gpurun_out/conv_table_pp$pp.json 2>gpurun_out/bench.err | tee gpurun_out/bench_pp$pp.json | python -c "$show"
done
echo "=== bench segformer"; timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --table gpurun_out/conv_table_sf.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf.json | python -c "$show"
echo "=== bench dofa"; timeout 900 python bench.py --workload dofa_base --steps 6 --warmup 3 --table gpurun_out/conv_table_dofa.json 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa.json | python -c "$show"
echo "=== reference arm (unetpp)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
tail -5 gpurun_out/bench.err
echo "=== ncu dram+time per launch (unetpp, eager)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1700 -c 900 --csv --log-file gpurun_out/launches_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_launch_bench.log 2>&1; tail -1 gpurun_out/ncu_launch_bench.log | cut -c1-200
echo "=== ncu full: conv_fwd / conv_wgrad"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 40 -c 4 -o gpurun_out/prof_r8_conv_fwd -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_full_fwd.log 2>&1; tail -1 gpurun_out/ncu_full_fwd.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 20 -c 4 -o gpurun_out/prof_r8_conv_wgrad -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_full_wgrad.log 2>&1; tail -1 gpurun_out/ncu_full_wgrad.log | cut -c1-120
ls -la gpurun_out | tail -20
