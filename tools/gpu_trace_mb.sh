#!/bin/bash
mkdir -p gpurun_out
for idx in 111 112 114; do
  BNB_MB_TRACE=gpurun_out/mbtrace_$idx.txt BNB_MB_TRACE_IDX=$idx timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 > /dev/null 2>&1
  echo "== mb trace $idx"; head -28 gpurun_out/mbtrace_$idx.txt
done
