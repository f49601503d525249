#!/bin/bash
mkdir -p gpurun_out
tr() { name=$1; shift; env "$@" BNB_MB2_TRACE=gpurun_out/trace_$name.txt BNB_MB2_TRACE_IDX=36 timeout 200 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 2>&1 | grep -E "^ +(36|40) " ; echo "-- $name"; head -9 gpurun_out/trace_$name.txt | cut -c1-120; }
tr base X=1
tr n48 BNB_MB2_NMMA=48
tr n112 BNB_MB2_NMMA=112
tr noload BNB_MB2_DBG=1
tr noloadstore BNB_MB2_DBG=5
tr pdl0 BNB_PDL=0
