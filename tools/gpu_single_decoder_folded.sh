#!/bin/bash
# single GPU: the folded SegFormer decoder (option "decoder_folded") — the whole -m gpu suite as the driver runs it, then
# SegFormer-B2 training and SegFormer-B5 sliding-window inference with the decoder folded (1) and in the reference's op order (0)
mkdir -p gpurun_out
echo "=== pytest -m gpu -x (as the driver runs it)"
SECONDS=0
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/rf_pytest_gpu_full.log 2>&1
tail -6 gpurun_out/rf_pytest_gpu_full.log; echo "wall=${SECONDS}s"
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "folded", d["config"].get("decoder_folded"), "launches", d["gpu_launches"], "roof", round(d["roofline"]["frac"],3), d["clocks"])'
for f in 1 0; do
  echo "=== segformer_b2 GDL_DECODER_FOLDED=$f"
  GDL_DECODER_FOLDED=$f timeout 300 python bench.py --workload segformer_b2 --steps 30 --warmup 3 --no-cpu-baseline --no-library-baseline --workloads headline 2>gpurun_out/fold.err | tee gpurun_out/rf_bench_sf_b2_folded$f.json | python -c "$show"
  tail -2 gpurun_out/fold.err
done
for f in 1 0; do
  echo "=== segformer_b5_infer raster 4096 GDL_DECODER_FOLDED=$f"
  GDL_DECODER_FOLDED=$f timeout 300 python bench.py --workload segformer_b5_infer --raster 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline 2>gpurun_out/fold.err | tee gpurun_out/rf_bench_infer_folded$f.json | python -c "$show"
  tail -2 gpurun_out/fold.err
done
