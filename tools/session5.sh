#!/usr/bin/env bash
# Multi-GPU session (gpurun --gpus N): sharded-engine tests, bench at N ranks with the parity check, A/B against the NCCL + library-barrier route, warm kernel times.
set -u
N=${1:-2}
T=${2:-s5}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PYTHONUNBUFFERED=1 timeout -k 10 420 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 200 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/${T}_tests.log | tail -12
timeout -k 10 240 $TR --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
echo "== bench N=$N rc=$?"; tail -c 1500 gpurun_out/${T}_bench_n$N.json; tail -5 gpurun_out/${T}_bench_n$N.err | cut -c1-300
RECHUB_B200_P2P_ALLREDUCE=0 RECHUB_B200_P2P_OWN_BARRIER=0 timeout -k 10 240 $TR --master-port 29532 bench.py --gpus $N --steps 100 --warmup 5 --no-check > gpurun_out/${T}_bench_n${N}_nccl.json 2> gpurun_out/${T}_bench_n${N}_nccl.err
echo "== bench (NCCL all-reduce + library barrier) rc=$?"; tail -c 400 gpurun_out/${T}_bench_n${N}_nccl.json | head -c 400; echo
timeout -k 10 300 $TR --master-port 29533 tools/kernel_times.py > gpurun_out/${T}_ktimes_n$N.txt 2>&1; head -34 gpurun_out/${T}_ktimes_n$N.txt | cut -c1-150
timeout -k 10 240 $TR --master-port 29534 bench.py --gpus $N --workload dcnv2 --steps 50 --warmup 5 > gpurun_out/${T}_bench_dcnv2_n$N.json 2> gpurun_out/${T}_bench_dcnv2_n$N.err
echo "== dcnv2 N=$N rc=$?"; tail -c 500 gpurun_out/${T}_bench_dcnv2_n$N.json; tail -3 gpurun_out/${T}_bench_dcnv2_n$N.err | cut -c1-300
timeout -k 10 240 $TR --master-port 29535 bench.py --gpus $N --workload din --steps 50 --warmup 5 > gpurun_out/${T}_bench_din_n$N.json 2> gpurun_out/${T}_bench_din_n$N.err
echo "== din N=$N rc=$?"; tail -c 500 gpurun_out/${T}_bench_din_n$N.json; tail -3 gpurun_out/${T}_bench_din_n$N.err | cut -c1-300
