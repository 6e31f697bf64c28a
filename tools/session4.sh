#!/usr/bin/env bash
# Full single-GPU session: whole GPU suite, default bench, warm kernel times, pipeline traces, batch sweep of the fused gather, configs 3-5.
set -u
mkdir -p gpurun_out
T=${1:-s4}
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${T}_tests.log | tail -25
timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
    print("BENCH value %.2f M/s  %.4f ms  e2e %.2f M/s  fwd %.2f us frac %.3f  gemm %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["avg_us"], d["roofline"]["frac"], d["roofline_gemm"]["us_per_step"]))
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/${T}_bench.err | cut -c1-300
timeout -k 10 200 python tools/kernel_times.py > gpurun_out/${T}_ktimes.txt 2>&1; head -30 gpurun_out/${T}_ktimes.txt | cut -c1-160
timeout -k 10 100 tools/gemm_trace > gpurun_out/${T}_gemm_trace.txt 2>&1
timeout -k 10 100 tools/bnfuse_trace > gpurun_out/${T}_bnfuse_trace.txt 2>&1; cat gpurun_out/${T}_bnfuse_trace.txt
timeout -k 10 300 python tools/sweep_fields_fwd.py > gpurun_out/${T}_sweep_fields_fwd.csv 2>&1; cat gpurun_out/${T}_sweep_fields_fwd.csv | tail -12
for w in dcnv2 din dssm; do
  timeout -k 10 400 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
  echo "== $w rc=$?"; tail -c 600 gpurun_out/${T}_bench_$w.json; tail -3 gpurun_out/${T}_bench_$w.err | cut -c1-400
done
# programmatic dependent launch: parity subset + bench A/B
RECHUB_B200_PDL=1 timeout -k 10 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_golden.py tests/test_gpu_bnfuse.py tests/test_gpu_gemm.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${T}_tests_pdl.log 2>&1
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
