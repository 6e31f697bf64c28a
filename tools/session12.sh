#!/usr/bin/env bash
# Round-2 kernel changes on one GPU: fused gather v6 (descriptors in shared memory), GEMM TMA-store epilogue + concatenated B,
# fills / opt_advance off the critical path.  Suite, traces, then bench A/B of each switch.
set -u
mkdir -p gpurun_out
T=${1:-s12}
PYTHONUNBUFFERED=1 timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/${T}_tests.log | tail -25
timeout -k 10 100 tools/fields_trace > gpurun_out/${T}_fields_trace_v6.txt 2>&1; echo "== fields_trace v6"; tail -16 gpurun_out/${T}_fields_trace_v6.txt
RECHUB_B200_FIELDS_FWD=4 timeout -k 10 100 tools/fields_trace > gpurun_out/${T}_fields_trace_v4.txt 2>&1; echo "== fields_trace v4"; tail -16 gpurun_out/${T}_fields_trace_v4.txt | head -8
timeout -k 10 100 tools/gemm_trace > gpurun_out/${T}_gemm_trace.txt 2>&1; echo "== gemm_trace (tma epilogue + concat b)"; grep -E "tile|exit|MMA: last|accumulator ready" gpurun_out/${T}_gemm_trace.txt | head -60
RECHUB_B200_GEMM_TMA_EPILOGUE=0 RECHUB_B200_GEMM_CONCAT_B=0 timeout -k 10 100 tools/gemm_trace > gpurun_out/${T}_gemm_trace_legacy.txt 2>&1; echo "== gemm_trace legacy"; grep -E "tile|exit|MMA: last|accumulator ready" gpurun_out/${T}_gemm_trace_legacy.txt | head -60
bench() {  # label env...
  local label=$1; shift
  env "$@" timeout -k 10 300 python bench.py --no-cpu-baseline --no-kernel-times > gpurun_out/${T}_bench_$label.json 2> gpurun_out/${T}_bench_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_$label.json").read().strip().splitlines()[-1])
    print("BENCH %-14s value %.2f M/s  %.4f ms  e2e %.2f M/s  fwd %.2f us frac %.3f  gemm %.1f us launches %s" % ("$label", d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["avg_us"], d["roofline"]["frac"], d["roofline_gemm"]["us_per_step"], d.get("gpu_launches_per_step")))
except Exception as e:
    print("bench $label failed", e)
PY
  tail -2 gpurun_out/${T}_bench_$label.err | cut -c1-200
}
bench default A=1
bench fwd_v4 RECHUB_B200_FIELDS_FWD=4
bench gemm_legacy RECHUB_B200_GEMM_TMA_EPILOGUE=0 RECHUB_B200_GEMM_CONCAT_B=0
bench gemm_epi_only RECHUB_B200_GEMM_CONCAT_B=0
bench late_advance RECHUB_B200_EARLY_OPT_ADVANCE=0
timeout -k 10 200 python tools/kernel_times.py > gpurun_out/${T}_warm_kernel_times.txt 2>&1; grep -v "Warn\|_warn_once" gpurun_out/${T}_warm_kernel_times.txt | head -24 | cut -c1-150
timeout -k 10 300 python tools/sweep_fields_fwd.py > gpurun_out/${T}_sweep_fields_fwd.csv 2>&1; tail -6 gpurun_out/${T}_sweep_fields_fwd.csv
