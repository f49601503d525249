#!/bin/bash
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/tma_box_bench.cu -o /tmp/tma_box_bench -lcuda && timeout 120 /tmp/tma_box_bench
