#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <timeout_s> <out_file> [--gpus N] -- <command>   — retries while the pod answers "busy" (exit 3)
T=$1; OUT=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > "$OUT" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$OUT"; then exit $rc; fi
  sleep 90
done
exit 3
