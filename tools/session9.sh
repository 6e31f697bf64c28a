#!/usr/bin/env bash
# GEMM tile-width check: tests, pipeline trace at both widths, bench A/B (auto vs forced 128).
set -u
mkdir -p gpurun_out
T=${1:-s9}
PYTHONUNBUFFERED=1 timeout -k 10 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_engine.py tests/test_gpu_golden.py tests/test_gpu_bnfuse.py tests/test_gpu_small_ops.py tests/test_gpu_parity.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" gpurun_out/${T}_tests.log | tail -25
timeout -k 10 100 tools/gemm_trace > gpurun_out/${T}_gemm_trace.txt 2>&1; grep -E "tile|exit|MMA: last|accumulator ready" gpurun_out/${T}_gemm_trace.txt
for tn in 0 128; do
  RECHUB_B200_GEMM_TILE_N=$tn timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_tn$tn.json 2> gpurun_out/${T}_bench_tn$tn.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_tn$tn.json").read().strip().splitlines()[-1])
    print("BENCH tile_n=$tn value %.2f M/s  %.4f ms  e2e %.2f M/s  fwd %.2f us  gemm %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["avg_us"], d["roofline_gemm"]["us_per_step"]))
except Exception as e:
    print("bench failed", e)
PY
done
timeout -k 10 200 python tools/kernel_times.py > gpurun_out/${T}_ktimes.txt 2>&1; grep -v Warn gpurun_out/${T}_ktimes.txt | head -24 | cut -c1-150
# software prefetch of the next batch's rows into L2 (e2e path), L2 fetch granularity, concurrent optimisers
for cfg in "RECHUB_B200_NEXT_BATCH_PREFETCH=1" "RECHUB_B200_L2_FETCH_GRANULARITY=64" "RECHUB_B200_CONCURRENT_OPT=1"; do
  env $cfg timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_ab.json 2> gpurun_out/${T}_bench_ab.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_ab.json").read().strip().splitlines()[-1])
    print("BENCH $cfg value %.2f M/s  %.4f ms  e2e %.2f M/s  fwd %.2f us" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["avg_us"]))
except Exception as e:
    print("bench $cfg failed", e)
PY
  tail -2 gpurun_out/${T}_bench_ab.err | cut -c1-300
done
