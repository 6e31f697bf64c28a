#!/usr/bin/env bash
# Round-2 single-GPU evidence (everything profiles/README.md indexes for r02): whole GPU suite, smoke, default bench (with cpu_baseline),
# other workloads, warm kernel times, ncu launch list of one step, ncu --set full of the hot kernels (summarised on the box), traces, sweeps.
set -u
mkdir -p gpurun_out
T=${1:-s10}
PYTHONUNBUFFERED=1 timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --durations=8 -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/${T}_tests.log | tail -25
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 400 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 2500 gpurun_out/${T}_bench_n1.json; tail -3 gpurun_out/${T}_bench_n1.err | cut -c1-300
timeout -k 10 200 python tools/kernel_times.py > gpurun_out/${T}_warm_kernel_times.txt 2>&1; grep -v Warn gpurun_out/${T}_warm_kernel_times.txt | head -32 | cut -c1-160
timeout -k 10 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_one_train_step.ncu.csv python tools/profile_step.py > gpurun_out/${T}_launches.log 2>&1
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"fields_fwd_v4|fields_bwd|bn_fused|gemm_tf32x3|rowwise_update" -o gpurun_out/${T}_prof python tools/profile_step.py > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
python tools/ncu_summary.py gpurun_out/${T}_prof.ncu-rep > gpurun_out/${T}_ncu_full_summary.json 2> gpurun_out/${T}_ncu_summary.err; head -c 600 gpurun_out/${T}_ncu_full_summary.json
timeout -k 10 100 tools/gemm_trace > gpurun_out/${T}_gemm_trace.txt 2>&1; grep -E "tile|exit|MMA: last|accumulator ready" gpurun_out/${T}_gemm_trace.txt | head -40
timeout -k 10 100 tools/bnfuse_trace > gpurun_out/${T}_bnfuse_trace.txt 2>&1; tail -12 gpurun_out/${T}_bnfuse_trace.txt
timeout -k 10 100 tools/fields_trace > gpurun_out/${T}_fields_trace.txt 2>&1; tail -30 gpurun_out/${T}_fields_trace.txt
timeout -k 10 200 tools/microbench_gather > gpurun_out/${T}_microbench_gather.csv 2>&1; tail -12 gpurun_out/${T}_microbench_gather.csv
timeout -k 10 300 python tools/sweep_fields_fwd.py > gpurun_out/${T}_sweep_fields_fwd.csv 2>&1; tail -8 gpurun_out/${T}_sweep_fields_fwd.csv
for w in dcnv2 din dssm; do
  timeout -k 10 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "== $w"; tail -c 1200 gpurun_out/${T}_bench_$w.json; tail -2 gpurun_out/${T}_bench_$w.err | cut -c1-300
done
timeout -k 10 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 800 gpurun_out/${T}_bench_reference.json
