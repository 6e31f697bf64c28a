#!/usr/bin/env bash
# 2-GPU check: the distributed GPU tests, then the scaling point at N=2 (parity + kernel_times inside the bench line).
set -u
mkdir -p gpurun_out
T=${1:-s11}
PYTHONUNBUFFERED=1 timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests_dist.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/${T}_tests_dist.log | tail -12
bash tools/session_n.sh 2 ${T}
