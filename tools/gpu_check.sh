#!/bin/bash
# One gpurun call: diagnostics first (they survive later failures), then tests, bench, profiles.
mkdir -p gpurun_out
MODE=$1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
echo "== layer report f32"; timeout 300 python tools/layer_report.py --precision f32 --out gpurun_out/layer_report_f32.txt 2>&1 | tail -70
echo "== layer report f16x3"; timeout 300 python tools/layer_report.py --precision f16x3 --out gpurun_out/layer_report_f16x3.txt 2>&1 | tail -70
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "== bench f16x3"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for cfg in "16 1" "16 2" "16 4" "32 1" "32 3" "64 1" "64 2" "128 1"; do set -- $cfg; echo "== bench f16x3 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
for cfg in "32 2" "64 2"; do set -- $cfg; echo "== bench f32 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --precision f32 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
if [ "$MODE" == "full" ]; then
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path.insert(0,'birdnet-go_b200')
import numpy as np, birdnet_b200 as bb
c=bb.B200Classifier(max_batch=3, micro_batch=2)
x=(0.1*np.random.default_rng(0).standard_normal((3,144000))).astype(np.float32)
print(c.analyze_batch(x)[0][:,0]); print(c.predict_batch((x*32767).astype(np.int16)).argmax(1))
" > gpurun_out/sanitizer.txt 2>&1; tail -8 gpurun_out/sanitizer.txt
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 64 > gpurun_out/ncu_bench.log 2>&1; tail -c 600 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches.csv
fi
