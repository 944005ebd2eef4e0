#!/bin/bash
# round 2, GPU call 7: fp32 split mode, ResNeXt parity, set_lr, ncu DRAM pass at the bench batch, ncu --set full of the wgrad kernels
mkdir -p gpurun_out
for f in tests/test_fp32_split_gpu.py tests/test_unetpp_gpu.py tests/test_zz1_inference_gpu.py tests/test_kernels_gpu.py; do
  b=$(basename "$f" .py)
  timeout 1500 python -m pytest "$f" -m gpu -q --no-header -rA -p no:cacheprovider > "gpurun_out/r7_$b.log" 2>&1
  echo "$b: $(grep -E ' passed| failed| error' "gpurun_out/r7_$b.log" | tail -1)"
  grep -E "^(FAILED|ERROR)|^E  " "gpurun_out/r7_$b.log" | head -12
done
grep -h "split3\|resnext" gpurun_out/r7_test_fp32_split_gpu.log gpurun_out/r7_test_unetpp_gpu.log | head -20
echo "=== bench resnext101_32x8d UNet++ (the shipped YAML's encoder), B=16"
timeout 900 python - <<'P' 2>&1 | tail -4
import sys, torch, time
sys.path.insert(0, 'geo-deep-learning_b200')
from gdl_b200 import ops
from gdl_b200.models.unetpp import UnetPlusPlus
from gdl_b200.trainer import FusedTrainer
torch.manual_seed(0)
B = 16
m = UnetPlusPlus("resnext101_32x8d", in_channels=3, classes=5).cuda().train()
tr = FusedTrainer(m, ops.LossSpec(1.0, 0.0, ignore_index=-100), lr=1e-4, mean=[0.5]*3, std=[0.2]*3, cuda_graph=True)
raw = torch.randint(0, 256, (B, 512, 512, 3), dtype=torch.uint8, device="cuda")
t = torch.randint(0, 5, (B, 512, 512), device="cuda")
for _ in range(3): l = tr.step(raw, t)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): l = tr.step(raw, t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"resnext101_32x8d UNet++ 3x512x512 B={B}: {ms:.1f} ms/step = {B / ms * 1e3:.1f} tiles/s, loss {l.item():.4f}, launches/step {tr.launches_per_step}")
P
echo "=== ncu: time + DRAM bytes per launch over one eager UNet++ step at the bench batch (B=32)"
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 7000 --csv \
  --log-file gpurun_out/r7_ncu_unetpp_b32_launches.csv python bench.py --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline --no-library-baseline --workloads headline > gpurun_out/r7_ncu_bench.log 2>&1
wc -l gpurun_out/r7_ncu_unetpp_b32_launches.csv
echo "=== ncu --set full: conv_wgrad_kernel / wgrad3x3_rows_kernel / conv_fwd_kernel (a few launches each, B=32 step)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'conv_wgrad_kernel|wgrad3x3_rows_kernel' --launch-skip 40 -c 8 -o gpurun_out/r7_wgrad_full \
  python bench.py --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline --no-library-baseline --workloads headline > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'conv_fwd_kernel|conv3x3_rows_kernel' --launch-skip 150 -c 8 -o gpurun_out/r7_fwd_full \
  python bench.py --steps 1 --warmup 1 --cuda-graph 0 --no-cpu-baseline --no-library-baseline --workloads headline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
