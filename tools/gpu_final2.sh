#!/bin/bash
# Round-2 end-of-round validation in one gpurun call: smoke, all GPU tests, host mirror, both bench arms, F32 truth path,
# compute-sanitizer memcheck + racecheck + synccheck (small shapes through the debug entry points and a 5-chunk classifier call).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
nproc > gpurun_out/r02_host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r02_host.txt; free -g | head -2 >> gpurun_out/r02_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02_host.txt 2>/dev/null
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.txt
echo "== host mirror (C++) on soundscape.wav"; (cd birdnet-go_b200/host && ./host_test analyze ../../assets/BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite ../../assets/BirdNET_GLOBAL_6K_V2.4_Labels_en_us.txt ../../assets/soundscape.wav 2>&1 | tail -4)
echo "== layer report f16x3"; timeout 300 python tools/layer_report.py --precision f16x3 --out gpurun_out/r02_layer_report_f16x3.txt 2>&1 | tail -3
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 1500 gpurun_out/r02_bench_reference.json
echo "== bench default"; timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 4500 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
echo "== bench f32"; timeout 300 python bench.py --steps 10 --warmup 3 --precision f32 --no-cpu-baseline --no-two-callers > gpurun_out/r02_bench_f32.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/r02_bench_f32.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']))"
echo "== bench config2 shape (batch 1024 synthetic)"; timeout 600 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline --no-two-callers > gpurun_out/r02_bench_config2.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/r02_bench_config2.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['config']['workload'][:80])"
SAN='
import sys; sys.path.insert(0,"birdnet-go_b200"); sys.path.insert(0,"tests")
import numpy as np, birdnet_b200 as bb
import os
n=int(os.environ.get("SAN_B","5"))
c=bb.B200Classifier(max_batch=n, micro_batch=2)
x=(0.1*np.random.default_rng(0).standard_normal((n,144000))).astype(np.float32)
print(c.analyze_batch(x)[0][:,0]); print(c.predict_batch((x*32767).astype(np.int16)).argmax(1))
print(c.analyze_batch_detections(x, 1.0, 0.01)[3])
'
for tool in memcheck racecheck synccheck; do
  if [ "$tool" == "memcheck" ]; then export SAN_B=5; else export SAN_B=2; fi
  echo "== sanitizer $tool (B=$SAN_B)"; timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -c "$SAN" > gpurun_out/r02_sanitizer_$tool.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_sanitizer_$tool.txt
done
