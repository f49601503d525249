#!/bin/bash
# kernel unit tests of the round-2 tensor-core kernels (each test resets the device after a CUDA error)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -150 | tee gpurun_out/k1_pytest.txt
