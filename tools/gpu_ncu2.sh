#!/bin/bash
# ncu --set full of the round-2 tensor-core kernels, one report per (kernel, layer)
mkdir -p gpurun_out
cap() { # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 --launch-count 1 -f -o gpurun_out/prof_$1 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu_$1.log 2>&1; tail -1 gpurun_out/ncu_$1.log
}
cap mb2_b1 mbconv2_kernel 16
cap mb2_b5 mbconv2_kernel 20
cap mb2_b9 mbconv2_kernel 24
cap mb2_b14 mbconv2_kernel 29
cap pw2_b1 pw2_kernel 18
cap pw2_b5 pw2_kernel 22
cap pw2_b9 pw2_kernel 26
ls -la gpurun_out/*.ncu-rep
echo "== launch times (resident b5-7)"; timeout 300 python tools/launch_times.py --micro-batch 64 --lanes 1 2>&1 | head -30
