#!/usr/bin/env bash
# Runs the host-executed CUDA kernel tests with AddressSanitizer + UBSan(alignment) on the kernel code:
# out-of-bounds global / shared-memory accesses and misaligned 128-bit accesses abort the run.
#   tools/hostemu_asan.sh [pytest args]        (default: tests/test_hostemu_kernels_cpu.py -q)
set -euo pipefail
cd "$(dirname "$0")/.."
export GDL_HOSTEMU_ASAN=1
export ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0:detect_stack_use_after_return=0
export LD_PRELOAD="$(gcc -print-file-name=libasan.so)"
if [ $# -eq 0 ]; then set -- tests/test_hostemu_kernels_cpu.py -q; fi
exec python -m pytest "$@"
