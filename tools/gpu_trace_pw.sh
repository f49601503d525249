#!/bin/bash
mkdir -p gpurun_out
for idx in 148 151 154; do
  BNB_PWTC_TRACE=gpurun_out/pwtrace_$idx.txt BNB_PWTC_TRACE_IDX=$idx timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 > /dev/null 2>&1
  echo "== pw trace $idx"; head -20 gpurun_out/pwtrace_$idx.txt
done
