#!/bin/bash
mkdir -p gpurun_out
for idx in 0 6; do
  BNB_PWTC_TRACE=gpurun_out/trace_$idx.txt BNB_PWTC_TRACE_IDX=$((246 + idx)) timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 > /dev/null 2>&1
  echo "== trace $idx"; head -30 gpurun_out/trace_$idx.txt
done
for idx in 73; do
  BNB_PWTC_TRACE=gpurun_out/trace_b$idx.txt BNB_PWTC_TRACE_IDX=$((246 + idx)) timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 > /dev/null 2>&1
  echo "== trace back $idx"; head -30 gpurun_out/trace_b$idx.txt
done
