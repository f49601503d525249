#!/bin/bash
mkdir -p gpurun_out
# fused launches per predict call: 4 micro x 8 front blocks + 8 back blocks = 40; 3 warm-up calls -> profiled call starts at 120
for idx in 4 32 37; do
  BNB_MB_TRACE=gpurun_out/mbtrace_$idx.txt BNB_MB_TRACE_IDX=$((120 + idx)) timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 > /dev/null 2>&1
  echo "== mb trace $idx"; head -40 gpurun_out/mbtrace_$idx.txt
done
