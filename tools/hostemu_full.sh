#!/usr/bin/env bash
# Reproduces profiles/r01_hostemu_full_run.txt: every GPU test on the host-executed CUDA sources (tests/hostemu), synchronous and
# asynchronous completion, AddressSanitizer, randomised shapes, whole models, mutation checks.  About an hour on 8 cores; no GPU.
set -uo pipefail
cd "$(dirname "$0")/.."
export GDL_HOSTEMU_FULL=1
echo "== 1. scalar kernels";            python -m pytest tests/test_hostemu_kernels_cpu.py -q -p no:cacheprovider | tail -1
echo "== 2. tensor-core kernels";       python -m pytest tests/test_hostemu_tensorcore_cpu.py -q -n 6 -p no:cacheprovider | tail -1
for seed in 2 3 4 5; do
  echo "== 2b. asynchronous completion, seed $seed"
  GDL_HOSTEMU_ASYNC=$seed GDL_HOSTEMU_FUZZ=16 python -m pytest tests/test_hostemu_tensorcore_cpu.py tests/test_hostemu_fuzz_cpu.py -q -n 6 \
    -k "not asynchronous_completion" -p no:cacheprovider | tail -1
done
echo "== 3. randomised shapes";         GDL_HOSTEMU_FUZZ=150 python -m pytest tests/test_hostemu_fuzz_cpu.py -q -n 6 -p no:cacheprovider | tail -1
echo "== 4. AddressSanitizer";          GDL_HOSTEMU_FUZZ=24 tools/hostemu_asan.sh tests/test_hostemu_kernels_cpu.py tests/test_hostemu_tensorcore_cpu.py \
                                          tests/test_hostemu_fuzz_cpu.py -q -n 3 -p no:cacheprovider | tail -1
echo "== 5. whole models, every kernel on the model"
GDL_HOSTEMU_TC=1 python -m pytest tests/test_hostemu_models_cpu.py -q -n 7 --timeout 3000 -p no:cacheprovider \
  -k "not dofa_unfrozen and not fused_trainer_reduces_loss and not test_training_reduces_loss" | tail -1
echo "== 6. mutation checks";           python tools/hostemu_mutation_check.py
echo "== 7. static counters";           python tools/hostemu_counters.py | tail -16
