#!/bin/bash
# round 2, run 2: kernel unit tests, end-to-end parity through the new plane path, launch times, bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== parity tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py tests/test_realtime_queue.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/k2_pytest.txt
echo "== launch times micro 64"; timeout 300 python tools/launch_times.py --micro-batch 64 --lanes 1 > gpurun_out/launch_times_v2_m64.txt 2>&1; cat gpurun_out/launch_times_v2_m64.txt | head -140
for cfg in "64 1" "64 2" "128 2"; do set -- $cfg; echo "== bench f16x3 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
echo "== bench v1 (BNB_V2=0) micro 64 lanes 2"; BNB_V2=0 timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch 64 --lanes 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
