#!/bin/bash
# run 17 (first GPU call of the next round): everything written after round 1's GPU budget was spent.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round17.sh'
mkdir -p gpurun_out
echo "=== new GPU tests (written blind): inference, augment/metrics, feeder, trainable DOFA, stochastic layers"
for f in tests/test_zz1_inference_gpu.py tests/test_zz2_augment_metrics_gpu.py tests/test_zz3_wds_feeder_gpu.py \
         tests/test_zz4_dofa_trainable_gpu.py tests/test_zz5_stochastic_layers_gpu.py tests/test_zz6_dynamic_encoder_gpu.py \
         tests/test_zz7_sra_attention_gpu.py; do
  timeout 900 python -m pytest "$f" -m gpu -q --no-header -rA -p no:cacheprovider > "gpurun_out/$(basename "$f" .py).log" 2>&1
  echo "$f: $(grep -E 'passed|failed|error' "gpurun_out/$(basename "$f" .py).log" | tail -1)"
  grep -E "^(FAILED|ERROR)|Error|assert " "gpurun_out/$(basename "$f" .py).log" | head -8
done
echo "=== full pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu_full.log | tail -2
echo "=== HBM-bound kernel microbench"
timeout 600 python tools/bench_hbm_kernels.py --out gpurun_out/hbm_kernels.json 2>gpurun_out/hbm.err | cut -c1-200
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["gpu_launches"])'
echo "=== bench default"; timeout 900 python bench.py --steps 8 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_unetpp.json | python -c "$show"
echo "=== bench dofa unfrozen"; timeout 900 python bench.py --workload dofa_base_unfrozen --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa_unfrozen.json | python -c "$show"
echo "=== bench infer 4096"; timeout 600 python bench.py --workload segformer_b5_infer --raster 4096 --steps 3 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_infer_4096.json | cut -c1-400
echo "=== fused SRA attention A/B (segformer_b2 training step, segformer_b5 sliding-window inference)"
for f in 0 1; do
  GDL_SRA_FUSED=$f timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf_b2_fused$f.json | python -c "$show"
  GDL_SRA_FUSED=$f timeout 600 python bench.py --workload segformer_b5_infer --raster 4096 --steps 3 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_infer_4096_fused$f.json | cut -c1-300
done
GDL_ATTN_WGRAD_GROUPED=1 timeout 600 python bench.py --workload segformer_b2 --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_sf_b2_wgrad_grouped.json | python -c "$show"
for f in 0 1; do
  GDL_MHA_FLASH=$f timeout 600 python bench.py --workload dofa_base --steps 8 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_dofa_flash$f.json | python -c "$show"
done
GDL_SRA_FUSED=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sra_attention -c 2 -o gpurun_out/sra_attention_full \
  python bench.py --workload segformer_b2 --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline > /dev/null 2>&1
echo "=== ncu: new kernels (time + dram bytes per launch)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/ncu_hbm_kernels.csv python tools/bench_hbm_kernels.py --iters 1 > /dev/null 2>&1
tail -3 gpurun_out/bench.err
