#!/usr/bin/env bash
# GPU session: full GPU suite, default bench, warm kernel times, GEMM pipeline trace, L2 fetch granularity seen by the engine.
set -u
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/s2_tests.log 2>&1
tail -40 gpurun_out/s2_tests.log
timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
tail -3 gpurun_out/s2_bench.err
timeout -k 10 200 python tools/kernel_times.py > gpurun_out/s2_ktimes.txt 2>&1
timeout -k 10 100 tools/gemm_trace > gpurun_out/s2_gemm_trace.txt 2>&1
timeout 100 python - > gpurun_out/s2_l2gran.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, "torch-rechub_b200")
import torch
from torch_rechub.b200 import _lib
_lib.err_flag(torch.device("cuda:0"))
print("l2 fetch granularity (before, now):", _lib.l2_fetch_granularity_seen)
PY
cat gpurun_out/s2_l2gran.txt
