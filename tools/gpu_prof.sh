#!/bin/bash
mkdir -p gpurun_out
echo "== launch times micro 64"; timeout 300 python tools/launch_times.py --micro-batch 64 --lanes 1 > gpurun_out/launch_times_m64.txt 2>&1; head -75 gpurun_out/launch_times_m64.txt
echo "== launch times micro 32 (front of list)"; timeout 300 python tools/launch_times.py --micro-batch 32 --lanes 1 > gpurun_out/launch_times_m32.txt 2>&1; head -3 gpurun_out/launch_times_m32.txt
for cfg in "32 1" "32 2" "64 1" "64 2" "128 1"; do set -- $cfg; echo "== bench f16x3 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['e2e'].get('h2d_gbs_measured'), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
echo "== bench f32 micro 64"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch 64 --lanes 1 --precision f32 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
echo "== ncu pw_tc"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc -c 5 -f -o gpurun_out/prof_pw python tools/launch_times.py --batch 64 --micro-batch 64 > gpurun_out/ncu_pw.log 2>&1; tail -2 gpurun_out/ncu_pw.log
echo "== ncu others"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"frontend_kernel|stem_mix|dw_conv" -c 4 -f -o gpurun_out/prof_misc python tools/launch_times.py --batch 64 --micro-batch 64 > gpurun_out/ncu_misc.log 2>&1; tail -2 gpurun_out/ncu_misc.log
ls -la gpurun_out/*.ncu-rep
