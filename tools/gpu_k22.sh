#!/bin/bash
mkdir -p gpurun_out
echo "== kernel + parity tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
b() { python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_synchronous_call']), d['latency_batch1_ms'], d['latency_batch1_ms_no_graph'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; }
export -f b
echo "== bench"; timeout 300 bash -c "b 128 2"
echo "== bench TEAMS=1"; BNB_PW2_TEAMS=1 timeout 300 bash -c "b 128 2"
