#!/usr/bin/env bash
# 2 GPUs: the distributed tests (fused hand-overs in both id layouts + the barrier route), then the scaling point with the hand-overs
# folded into the launches vs as barrier kernels (parity + kernel_times inside each bench line).
set -u
mkdir -p gpurun_out
T=${1:-s14}
PYTHONUNBUFFERED=1 timeout -k 10 700 python -m pytest tests/test_gpu_dist.py tests/test_gpu_gemm.py::test_tma_epilogue_stays_inside_the_16_byte_rows_of_c tests/test_gpu_engine.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests_dist.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/${T}_tests_dist.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cfg in fused barriers; do
  v=1; [ $cfg = barriers ] && v=0
  RECHUB_B200_P2P_FUSED_SYNC=$v timeout -k 10 300 $TR --master-port 2954$v bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/${T}_bench_n2_$cfg.json 2> gpurun_out/${T}_bench_n2_$cfg.err
  echo "== bench N=2 $cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_n2_$cfg.json").read().strip().splitlines()[-1])
    print("BENCH N=2 %-9s value %.2f M/s  %.4f ms  e2e %.2f M/s  parity %s  launches %s" % ("$cfg", d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["parity"].get("ok"), d.get("gpu_launches_per_step")))
    for k in d["kernel_times"]["top"][:16]: print("   %5.1f/step %7.2f us  %s" % (k["per_step"], k["avg_us"], k["kernel"][:90]))
except Exception as e:
    print("bench failed", e)
PY
  tail -3 gpurun_out/${T}_bench_n2_$cfg.err | cut -c1-300
done
