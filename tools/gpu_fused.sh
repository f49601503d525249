#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for cfg in "64 2" "96 2" "128 2" "64 3"; do set -- $cfg; echo "== bench f16x3 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
echo "== ncu pw_tc project of block 5"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_pwproj python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu_pwproj.log 2>&1; tail -2 gpurun_out/ncu_pwproj.log
