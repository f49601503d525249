#!/bin/bash
mkdir -p gpurun_out
for p in 2 1; do echo "== kernel + parity tests POLY=$p"; BNB_MB2_POLY=$p timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4; done
b() { python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_synchronous_call']), d['latency_batch1_ms'], d['latency_batch1_ms_no_graph'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; }
export -f b
for p in 0 1 2; do echo "== bench POLY=$p"; BNB_MB2_POLY=$p timeout 300 bash -c "b 128 2"; done
echo "== layer report POLY=2"; BNB_MB2_POLY=2 timeout 300 python tools/layer_report.py --precision f16x3 --out gpurun_out/layer_report_poly2.txt 2>&1 | tail -2
