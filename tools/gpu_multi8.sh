#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
echo "== 8-GPU bench (default flags, as the driver launches it)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; python -c "import json; d=json.loads(open('gpurun_out/r02_bench_8gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"; tail -2 gpurun_out/r02_bench_8gpu.err
