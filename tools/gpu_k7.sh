#!/bin/bash
# MMA issue-order experiments: units interleaved per round (mbconv2) x accumulators per tile (pw2)
mkdir -p gpurun_out
for cfg in "1 1" "1 3" "2 3" "4 1" "2 2"; do set -- $cfg
  echo "== BNB_MB2_BATCH=$1 BNB_PW2_NACC=$2"
  BNB_MB2_BATCH=$1 BNB_PW2_NACC=$2 timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch 64 --lanes 2 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_synchronous_call']), d['latency_batch1_ms'], d['latency_batch1_ms_no_graph'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
done
echo "== new tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py tests/test_range_filter.py -m gpu -q -p no:cacheprovider -k "benched or config3 or async or nan_and or two_devices or batched_offline or range_filter" 2>&1 | tail -15
BNB_MB2_BATCH=1 BNB_MB2_TRACE=gpurun_out/trace_mb2_b5_w1.txt BNB_MB2_TRACE_IDX=36 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1; head -16 gpurun_out/trace_mb2_b5_w1.txt
BNB_MB2_BATCH=2 BNB_MB2_TRACE=gpurun_out/trace_mb2_b5_w2.txt BNB_MB2_TRACE_IDX=36 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1; head -16 gpurun_out/trace_mb2_b5_w2.txt
