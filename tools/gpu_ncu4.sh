#!/bin/bash
# refresh of the pw2 captures after the 256-bit epilogue stores (final tree)
mkdir -p gpurun_out
cap() { timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$2 --launch-skip $3 --launch-count 1 -f -o gpurun_out/prof4_$1 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu4_$1.log 2>&1; tail -1 gpurun_out/ncu4_$1.log; }
cap pw2_blk1 pw2_kernel 19
cap pw2_blk4 pw2_kernel 22
ls -la gpurun_out/prof4_*.ncu-rep
