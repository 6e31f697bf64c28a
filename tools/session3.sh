#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for cfg in "X=1" "RECHUB_B200_FUSED_BN=0 RECHUB_B200_FUSED_BN_HEAD=0" "RECHUB_B200_CONCURRENT_BWD=0" "RECHUB_B200_TC_GEMM=0" "DIAG_HEAD=0"; do
  echo "=== $cfg"; env $cfg timeout 120 python tools/diag_tower.py 2>&1 | grep -v Warning
done > gpurun_out/s3_diag.txt 2>&1
cat gpurun_out/s3_diag.txt
timeout -k 10 100 tools/gemm_trace > gpurun_out/s3_gemm_trace.txt 2>&1
timeout -k 10 100 tools/bnfuse_trace > gpurun_out/s3_bnfuse_trace.txt 2>&1; cat gpurun_out/s3_bnfuse_trace.txt
timeout -k 10 600 python -m pytest tests/test_gpu_bnfuse.py tests/test_gpu_crossmix.py tests/test_gpu_gemm.py tests/test_gpu_engine.py tests/test_gpu_golden.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/s3_tests.log 2>&1; tail -15 gpurun_out/s3_tests.log
timeout -k 10 300 python bench.py --no-cpu-baseline > gpurun_out/s3_bench.json 2> /dev/null
PROF_GRAPH=0 PROF_STEPS=3 timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:"fields_fwd_v5|bn_fused_fwd" -s 8 -c 6 -o gpurun_out/s3_prof python tools/kernel_times.py > gpurun_out/s3_ncu.log 2>&1; tail -3 gpurun_out/s3_ncu.log
