#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_rate tools/micro/mma_rate.cu -I rgb_no_more_b200/csrc 2>&1 | tail -3
timeout 60 /tmp/mma_rate | tee gpurun_out/mma_rate.log
