#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6
echo "== parity tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
for cfg in "64 2" "128 2"; do set -- $cfg; echo "== bench f16x3 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_synchronous_call']), d['latency_batch1_ms'], d['latency_batch1_ms_no_graph'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
tr() { if [ "$1" == "mb2" ]; then BNB_MB2_TRACE=gpurun_out/trace_$3.txt BNB_MB2_TRACE_IDX=$2 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1
  else BNB_PW2_TRACE=gpurun_out/trace_$3.txt BNB_PW2_TRACE_IDX=$2 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1; fi; echo "-- $3"; head -20 gpurun_out/trace_$3.txt; }
tr mb2 36 mb2_b5
tr mb2 33 mb2_b2
tr pw2 36 pw2_b1
tr pw2 40 pw2_b5
