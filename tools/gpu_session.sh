#!/usr/bin/env bash
# One GPU session over everything that was written without a GPU behind it (run from the repo root on a B200):
#   gpurun --timeout 900 -- 'bash tools/gpu_session.sh > gpurun_out/session.log 2>&1; tail -40 gpurun_out/session.log'
# Each stage is bounded by its own timeout; a failing stage does not stop the following ones.
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; "$@"; echo "=== rc=$?"; }
run timeout 300 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider
run env RECHUB_B200_FUSED_HEAD_ALL=1 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q --timeout 120 -p no:cacheprovider
run timeout 200 python bench.py --no-cpu-baseline
run env RECHUB_B200_NEXT_BATCH_PREFETCH=1 timeout 200 python bench.py --no-cpu-baseline
run timeout 200 python bench.py --no-cpu-baseline --ids zipf
run timeout 120 python tools/kernel_times.py
run timeout 200 python tools/sweep_gemm.py
