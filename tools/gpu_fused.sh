#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for bsz in 256 128 64; do echo "== launch times batch $bsz micro 64"; timeout 300 python tools/launch_times.py --batch $bsz --micro-batch 64 --lanes 1 > gpurun_out/launch_times_b$bsz.txt 2>&1; head -1 gpurun_out/launch_times_b$bsz.txt; tail -36 gpurun_out/launch_times_b$bsz.txt | head -20; done
for r in 0 1 2; do echo "== bench BNB_RAMP=$r"; BNB_RAMP=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-two-callers 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_int16_pcm']))"; done
