#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6
echo "== parity tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12
b() { python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_synchronous_call']), d['latency_batch1_ms'], d['latency_batch1_ms_no_graph'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; }
export -f b
for pdl in 1 0; do for cfg in "64 2" "128 2" "64 1" "64 3"; do set -- $cfg; echo "== bench PDL=$pdl micro $1 lanes $2"; BNB_PDL=$pdl timeout 300 bash -c "b $1 $2"; done; done
echo "== bench NACC=3 micro 64 lanes 2"; BNB_PW2_NACC=3 timeout 300 bash -c "b 64 2"
tr() { if [ "$1" == "mb2" ]; then BNB_MB2_TRACE=gpurun_out/trace_$3.txt BNB_MB2_TRACE_IDX=$2 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1
  else BNB_PW2_TRACE=gpurun_out/trace_$3.txt BNB_PW2_TRACE_IDX=$2 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1; fi; echo "-- $3"; head -16 gpurun_out/trace_$3.txt; }
tr mb2 36 mb2_b5
tr pw2 40 pw2_b5
echo "== launch times"; BNB_PDL=0 timeout 200 python tools/launch_times.py --batch 256 --micro-batch 64 --lanes 1 2>&1 | tee gpurun_out/launch_times_v5_m64.txt | head -30
