#!/bin/bash
mkdir -p gpurun_out
# launch order per front micro-batch: E1 P1 E2 P2 E3 P3 E4 P4 E5 P5 E6 P6 E7 P7 E8 P8 = 16 pw launches; warm-up 3 calls x (4 micro x 16 + 18 back)
for idx in 0 1 6 15; do
  BNB_PWTC_TRACE=gpurun_out/trace_$idx.txt BNB_PWTC_TRACE_IDX=$((246 + idx)) timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 > /dev/null 2>&1
  echo "== trace $idx"; head -14 gpurun_out/trace_$idx.txt
done
# back phase: after 4*16 front launches: P layers of S4 etc.
for idx in 64 65 72 73 80 81; do
  BNB_PWTC_TRACE=gpurun_out/trace_b$idx.txt BNB_PWTC_TRACE_IDX=$((246 + idx)) timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 > /dev/null 2>&1
  echo "== trace back $idx"; head -12 gpurun_out/trace_b$idx.txt
done
