#!/bin/bash
# round 2, GPU call 1: adjudicate graph-vs-eager, run EVERY -m gpu test (no -x), then the round-17 measurement list.
mkdir -p gpurun_out
echo "=== diag graph vs eager"
timeout 300 python tools/diag_graph_vs_eager.py --hw 64 --batch 4 2>gpurun_out/diag.err | tee gpurun_out/diag_64.json | cut -c1-1500
timeout 300 python tools/diag_graph_vs_eager.py --hw 128 --batch 8 2>>gpurun_out/diag.err | tee gpurun_out/diag_128.json | cut -c1-1500
tail -3 gpurun_out/diag.err
echo "=== per-file pytest -m gpu"
for f in tests/test_*_gpu.py; do
  b=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --no-header -rA -p no:cacheprovider > "gpurun_out/$b.log" 2>&1
  echo "$b: $(grep -E ' passed| failed| error' "gpurun_out/$b.log" | tail -1)"
  grep -E "^(FAILED|ERROR)" "gpurun_out/$b.log" | head -12
done
echo "=== HBM-bound kernel microbench"
timeout 600 python tools/bench_hbm_kernels.py --out gpurun_out/hbm_kernels.json 2>gpurun_out/hbm.err | cut -c1-200
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["gpu_launches"])'
echo "=== bench default"; timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_unetpp.json | python -c "$show"
echo "=== bench dofa unfrozen"; timeout 900 python bench.py --workload dofa_base_unfrozen --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa_unfrozen.json | python -c "$show"
echo "=== bench infer 4096"; timeout 600 python bench.py --workload segformer_b5_infer --raster 4096 --steps 3 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_infer_4096.json | cut -c1-400
echo "=== fused SRA attention A/B"
for f in 0 1; do
  GDL_SRA_FUSED=$f timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf_b2_fused$f.json | python -c "$show"
  GDL_SRA_FUSED=$f timeout 600 python bench.py --workload segformer_b5_infer --raster 4096 --steps 3 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_infer_4096_fused$f.json | cut -c1-300
done
GDL_ATTN_WGRAD_GROUPED=1 timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf_b2_wgrad_grouped.json | python -c "$show"
for f in 0 1; do
  GDL_MHA_FLASH=$f timeout 600 python bench.py --workload dofa_base --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa_flash$f.json | python -c "$show"
done
GDL_SRA_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sra_attention -c 3 -o gpurun_out/sra_attention_full \
  python bench.py --workload segformer_b2 --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline > gpurun_out/ncu_sra.log 2>&1
echo "=== ncu launch list segformer_b2 (fused)"
GDL_SRA_FUSED=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv \
  --log-file gpurun_out/ncu_sf_b2_launches.csv python bench.py --workload segformer_b2 --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline > /dev/null 2>&1
tail -5 gpurun_out/bench.err
