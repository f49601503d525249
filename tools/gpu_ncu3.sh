#!/bin/bash
# r02 evidence: ncu --set full of the F16X3 kernels (one report per kernel / layer), the launch list of one real bench step
# (two lanes, --cache-control none: DRAM bytes as the schedule really moves them), SASS opcode counts.
mkdir -p gpurun_out
cap() { # name regex skip
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$2 --launch-skip $3 --launch-count 1 -f -o gpurun_out/prof3_$1 python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu3_$1.log 2>&1; tail -1 gpurun_out/ncu3_$1.log
}
cap mb2_blk4 mbconv2_kernel 20
cap mb2_blk9 mbconv2_kernel 25
cap pw2_blk1 pw2_kernel 19
cap pw2_blk4 pw2_kernel 22
cap pw2_blk9 pw2_kernel 27
cap frontend frontend_kernel 1
cap stem stem_mix_kernel 1
ls -la gpurun_out/prof3_*.ncu-rep
echo "== launch list of one bench step (2 lanes, micro 128, cache-control none)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/launches3_dram.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-two-callers > gpurun_out/ncu3_bench.log 2>&1; tail -c 200 gpurun_out/ncu3_bench.log; wc -l gpurun_out/launches3_dram.csv
