#!/bin/bash
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['config'], d['cpu_baseline']['value'], sorted(d.keys()))"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
