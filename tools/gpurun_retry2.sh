#!/bin/bash
# usage: tools/gpurun_retry2.sh <logfile> <timeout_s> <n_gpus> <command...>  -- like gpurun_retry.sh with --gpus N (charged N x)
log=$1; to=$2; n=$3; shift 3
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$to" -- "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then echo "exit $rc" >> "$log"; exit $rc; fi
  sleep 60
done
echo "exit 3 (gave up)" >> "$log"; exit 3
