#!/usr/bin/env bash
# Whole GPU suite on the round-2 kernels (GEMM TMA epilogue + concatenated B, dW targets zeroed in the forward), bench, warm kernel times,
# ncu --set full of the GEMM / BatchNorm / gather kernels of one step (summarised on the box).
set -u
mkdir -p gpurun_out
T=${1:-s13}
PYTHONUNBUFFERED=1 timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --durations=6 -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/${T}_tests.log | tail -25
bench() {
  local label=$1; shift
  env "$@" timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_$label.json 2> gpurun_out/${T}_bench_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_$label.json").read().strip().splitlines()[-1])
    print("BENCH %-14s value %.2f M/s  %.4f ms  e2e %.2f M/s  fwd %.2f us frac %.3f  gemm %.1f us launches %s loss %.5f" % ("$label", d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["avg_us"], d["roofline"]["frac"], d["roofline_gemm"]["us_per_step"], d.get("gpu_launches_per_step"), d["final_loss"]))
except Exception as e:
    print("bench $label failed", e)
PY
  tail -2 gpurun_out/${T}_bench_$label.err | cut -c1-200
}
bench default A=1
bench no_prezero RECHUB_B200_PREZERO_DW=0
timeout -k 10 200 python tools/kernel_times.py > gpurun_out/${T}_warm_kernel_times.txt 2>&1; grep -v "Warn\|_warn_once" gpurun_out/${T}_warm_kernel_times.txt | head -26 | cut -c1-150
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"fields_fwd_v4|fields_bwd|bn_fused|gemm_tf32x3|rowwise_update" -o gpurun_out/${T}_prof python tools/profile_step.py > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
python tools/ncu_summary.py gpurun_out/${T}_prof.ncu-rep > gpurun_out/${T}_ncu_full_summary.json 2> gpurun_out/${T}_ncu_summary.err; head -c 300 gpurun_out/${T}_ncu_full_summary.json
rm -f gpurun_out/${T}_prof.ncu-rep
timeout -k 10 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_one_train_step.ncu.csv python tools/profile_step.py > gpurun_out/${T}_launches.log 2>&1
