#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do echo "== dbg $d"; BNB_PWTC_DBG=$d timeout 200 python tools/launch_times.py --micro-batch 64 --lanes 1 2>&1 | sed -n '1p;5p;7p;8p;14p;18p;30p;33p'; done
