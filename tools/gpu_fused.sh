#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== launch times resident weights"; timeout 300 python tools/launch_times.py --micro-batch 64 --lanes 1 > gpurun_out/launch_times_m64.txt 2>&1; head -26 gpurun_out/launch_times_m64.txt; tail -42 gpurun_out/launch_times_m64.txt | head -40
echo "== launch times BNB_PWTC_BRES=0"; BNB_PWTC_BRES=0 timeout 300 python tools/launch_times.py --micro-batch 64 --lanes 1 > gpurun_out/launch_times_m64_nores.txt 2>&1; head -1 gpurun_out/launch_times_m64_nores.txt
echo "== bench"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-two-callers 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
