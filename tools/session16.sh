#!/usr/bin/env bash
# Which kernel families may be launched early?  The two graph-replay tests that fail with every family under PDL, per family mask;
# then the union of the passing families: tests + bench against the no-PDL baseline.
set -u
mkdir -p gpurun_out
T=${1:-s16}
TESTS="tests/test_gpu_engine.py::test_cuda_graph_replay_equals_eager_steps tests/test_gpu_fullshape.py::test_deepfm_benchmarked_step_graph_replayed_rowwise_adam_against_oracle"
ok=0
for m in 1 2 4 8 16 32; do
  RECHUB_B200_PDL=$m PYTHONUNBUFFERED=1 timeout -k 10 200 python -m pytest $TESTS -m gpu -q --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests_m$m.log 2>&1
  r=$(grep -E "passed|failed" gpurun_out/${T}_tests_m$m.log | tail -1)
  echo "mask $m: $r"
  if echo "$r" | grep -q "2 passed"; then ok=$((ok | m)); fi
done
echo "passing families: mask $ok"
bench() {
  local label=$1; shift
  env "$@" timeout -k 10 300 python bench.py --no-cpu-baseline --no-kernel-times > gpurun_out/${T}_bench_$label.json 2> gpurun_out/${T}_bench_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_$label.json").read().strip().splitlines()[-1])
    print("BENCH %-10s value %.2f M/s  %.4f ms  e2e %.2f M/s  loss %.5f" % ("$label", d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["final_loss"]))
except Exception as e:
    print("bench $label failed", e)
PY
}
bench pdl0 RECHUB_B200_PDL=0
if [ $ok -ne 0 ]; then
  RECHUB_B200_PDL=$ok PYTHONUNBUFFERED=1 timeout -k 10 300 python -m pytest $TESTS tests/test_gpu_golden.py tests/test_gpu_bnfuse.py -m gpu -q --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests_union.log 2>&1
  echo "union mask $ok: $(grep -E 'passed|failed' gpurun_out/${T}_tests_union.log | tail -1)"
  bench pdl_union RECHUB_B200_PDL=$ok
fi
