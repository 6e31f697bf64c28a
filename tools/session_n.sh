#!/usr/bin/env bash
# Scaling point at N ranks: bench line (parity + kernel_times inside) and nothing else.  usage: session_n.sh N TAG
set -u
N=${1:-8}
T=${2:-sN}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29541 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
echo "== bench N=$N rc=$?"; tail -c 3000 gpurun_out/${T}_bench_n$N.json; tail -5 gpurun_out/${T}_bench_n$N.err | cut -c1-300
