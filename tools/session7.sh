#!/usr/bin/env bash
# Single-GPU check after the barrier fix: suite (thread-method timeouts so a hung kernel ends the run), bench, kernel times, traces, PDL A/B.
set -u
mkdir -p gpurun_out
T=${1:-s7}
PYTHONUNBUFFERED=1 timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method=thread --durations=12 -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" gpurun_out/${T}_tests.log | tail -25
timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
    print("BENCH value %.2f M/s  %.4f ms  e2e %.2f M/s  fwd %.2f us frac %.3f  gemm %.1f us launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["avg_us"], d["roofline"]["frac"], d["roofline_gemm"]["us_per_step"], d.get("gpu_launches")))
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/${T}_bench.err | cut -c1-300
timeout -k 10 200 python tools/kernel_times.py > gpurun_out/${T}_ktimes.txt 2>&1; grep -v Warn gpurun_out/${T}_ktimes.txt | head -28 | cut -c1-150
timeout -k 10 100 tools/fields_trace > gpurun_out/${T}_fields_trace.txt 2>&1; cat gpurun_out/${T}_fields_trace.txt
timeout -k 10 100 tools/bnfuse_trace > gpurun_out/${T}_bnfuse_trace.txt 2>&1; cat gpurun_out/${T}_bnfuse_trace.txt
RECHUB_B200_PDL=1 PYTHONUNBUFFERED=1 timeout -k 10 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_golden.py tests/test_gpu_bnfuse.py tests/test_gpu_gemm.py tests/test_gpu_small_ops.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests_pdl.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${T}_tests_pdl.log | tail -8
RECHUB_B200_PDL=1 timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_pdl.json 2> gpurun_out/${T}_bench_pdl.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_pdl.json").read().strip().splitlines()[-1])
    print("BENCH PDL value %.2f M/s  %.4f ms  e2e %.2f M/s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("bench pdl failed", e)
PY
tail -3 gpurun_out/${T}_bench_pdl.err | cut -c1-300
