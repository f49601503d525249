#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== launch times micro 64"; timeout 300 python tools/launch_times.py --micro-batch 64 --lanes 1 > gpurun_out/launch_times_m64.txt 2>&1; head -1 gpurun_out/launch_times_m64.txt; tail -42 gpurun_out/launch_times_m64.txt | head -40
for cfg in "64 1" "64 2"; do set -- $cfg; echo "== bench f16x3 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
