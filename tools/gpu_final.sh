#!/bin/bash
# Round-end validation in one gpurun call: diagnostics, parity reports, smoke, GPU tests, both bench arms, host mirror,
# compute-sanitizer, ncu launch list (time + DRAM bytes) of the bench command, ncu --set full of the two tcgen05 kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
echo "== layer report f32"; timeout 300 python tools/layer_report.py --precision f32 --out gpurun_out/layer_report_f32.txt 2>&1 | tail -4
echo "== layer report f16x3"; timeout 300 python tools/layer_report.py --precision f16x3 --out gpurun_out/layer_report_f16x3.txt 2>&1 | tail -4
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== host mirror (C++) on soundscape.wav"; (cd birdnet-go_b200/host && ./host_test analyze ../../assets/BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite ../../assets/BirdNET_GLOBAL_6K_V2.4_Labels_en_us.txt ../../assets/soundscape.wav 2>&1 | tail -4)
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1200 gpurun_out/bench_reference.json
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench f32"; timeout 300 python bench.py --steps 10 --warmup 3 --precision f32 --no-cpu-baseline > gpurun_out/bench_f32.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/bench_f32.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']))"
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path.insert(0,'birdnet-go_b200')
import numpy as np, birdnet_b200 as bb
c=bb.B200Classifier(max_batch=5, micro_batch=2)
x=(0.1*np.random.default_rng(0).standard_normal((5,144000))).astype(np.float32)
print(c.analyze_batch(x)[0][:,0]); print(c.predict_batch((x*32767).astype(np.int16)).argmax(1))
" > gpurun_out/sanitizer.txt 2>&1; tail -4 gpurun_out/sanitizer.txt
echo "== ncu launch list (time + dram bytes) of bench.py"; timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_dram.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -c 300 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_dram.csv
echo "== ncu full: mbconv_tc (block 5) and pw_tc (block 5 project)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mbconv_tc --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_mb python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu_mb.log 2>&1; tail -1 gpurun_out/ncu_mb.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_pwproj python tools/launch_times.py --batch 64 --micro-batch 64 --lanes 1 > gpurun_out/ncu_pwproj.log 2>&1; tail -1 gpurun_out/ncu_pwproj.log
