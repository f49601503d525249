#!/bin/bash
mkdir -p gpurun_out
echo "== ncu mbconv_tc (launch 5 = block 5, 12x32 map)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mbconv_tc --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_mb python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu_mb.log 2>&1; tail -2 gpurun_out/ncu_mb.log
echo "== ncu frontend + stem"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"frontend_kernel|stem_mix" -c 2 -f -o gpurun_out/prof_fe python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu_fe.log 2>&1; tail -2 gpurun_out/ncu_fe.log
ls -la gpurun_out/*.ncu-rep
