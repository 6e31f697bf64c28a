#!/usr/bin/env bash
# Programmatic dependent launch WITH the early trigger (griddepcontrol.launch_dependents at the top of every hot-path kernel): suite
# under PDL, then bench A/B on one box.
set -u
mkdir -p gpurun_out
T=${1:-s15}
RECHUB_B200_PDL=1 PYTHONUNBUFFERED=1 timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider --deselect tests/test_gpu_gemm.py::test_kernel_variants > gpurun_out/${T}_tests_pdl.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/${T}_tests_pdl.log | tail -12
bench() {
  local label=$1; shift
  env "$@" timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_$label.json 2> gpurun_out/${T}_bench_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_$label.json").read().strip().splitlines()[-1])
    print("BENCH %-10s value %.2f M/s  %.4f ms  e2e %.2f M/s  fwd %.2f us  gemm %.1f us  loss %.5f  kernel sum %.1f us" % ("$label", d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["avg_us"], d["roofline_gemm"]["us_per_step"], d["final_loss"], d["kernel_times"]["sum_us_per_step"]))
except Exception as e:
    print("bench $label failed", e)
PY
  tail -2 gpurun_out/${T}_bench_$label.err | cut -c1-200
}
bench pdl0 RECHUB_B200_PDL=0
bench pdl1 RECHUB_B200_PDL=1
bench pdl0b RECHUB_B200_PDL=0
bench pdl1b RECHUB_B200_PDL=1
for w in dcnv2 din; do
  for v in 0 1; do
    RECHUB_B200_PDL=$v timeout -k 10 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${T}_bench_${w}_pdl$v.json 2> gpurun_out/${T}_bench_${w}_pdl$v.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_${w}_pdl$v.json").read().strip().splitlines()[-1])
    print("BENCH $w pdl=$v value %.3f M/s  %.4f ms  e2e %.3f M/s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("bench $w failed", e)
PY
  done
done
