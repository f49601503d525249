#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "== layer report f16x3"; timeout 300 python tools/layer_report.py --precision f16x3 --out gpurun_out/layer_report_k17.txt 2>&1 | tail -3; head -12 gpurun_out/layer_report_k17.txt
b() { python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_synchronous_call']), d['latency_batch1_ms'], d['latency_batch1_ms_no_graph'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; }
export -f b
echo "== bench micro 128 lanes 2"; timeout 300 bash -c "b 128 2"
