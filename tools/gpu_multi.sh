#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
echo "== 2-GPU bench"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --micro-batch 64 --lanes 2 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1800 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
echo "== 1-GPU bench same config"; timeout 600 python bench.py --steps 10 --warmup 3 --micro-batch 64 --lanes 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
for cfg in "32 3" "32 4" "64 3" "16 4"; do set -- $cfg; echo "== bench micro $1 lanes $2"; timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $1 --lanes $2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; done
