#!/bin/bash
# 2-GPU checks (gpurun --gpus 2): the driver's launch line, config-2 shape sharded over two ranks, the reference arm under torchrun,
# the tests that need two devices or torch.distributed.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
echo "== 2-GPU bench (default flags, as the driver launches it)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -c 900 gpurun_out/r02_bench_2gpu.json | head -c 900; echo; tail -3 gpurun_out/r02_bench_2gpu.err
echo "== 1-GPU bench same box"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-two-callers 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"
echo "== 2-GPU config2 shape (2048 synthetic chunks per step, contiguous shards)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload config2 --steps 5 --warmup 3 --no-two-callers > gpurun_out/r02_bench_2gpu_config2.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/r02_bench_2gpu_config2.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['config'])"
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
echo "== gpu tests that use several devices / torch.distributed"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "dist or shard or realtime or two_devices" 2>&1 | tail -3
