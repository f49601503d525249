#!/bin/bash
b() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; }
export -f b
echo "== default"; timeout 300 bash -c "b"
echo "== TEAMS=2 everywhere"; BNB_PW2_TEAMS=2 timeout 300 bash -c "b"
echo "== lanes 3"; timeout 300 python bench.py --steps 10 --warmup 3 --lanes 3 --no-cpu-baseline --no-two-callers 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"
