#!/bin/bash
# round 2, run 4: determinism re-check, clock64 timelines of mbconv2 / pw2, bench variants
mkdir -p gpurun_out
echo "== invariance tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "int16 or composition or linearity or tensor_core_path or 256_invariants" 2>&1 | tail -8
tr() { # kind idx name
  if [ "$1" == "mb2" ]; then BNB_MB2_TRACE=gpurun_out/trace_$3.txt BNB_MB2_TRACE_IDX=$2 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1
  else BNB_PW2_TRACE=gpurun_out/trace_$3.txt BNB_PW2_TRACE_IDX=$2 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > /dev/null 2>&1; fi
  echo "-- $3"; cat gpurun_out/trace_$3.txt
}
tr mb2 32 mb2_b1
tr mb2 36 mb2_b5
tr mb2 40 mb2_b9
tr mb2 45 mb2_b14
tr pw2 36 pw2_b1
tr pw2 40 pw2_b5
tr pw2 44 pw2_b9
for cfg in "64 2" "64 4" "32 4" "128 2"; do set -- $cfg; echo "== bench f16x3 micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['latency_batch1_ms'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; done
