#!/bin/bash
mkdir -p gpurun_out
./tools/micro/readbw > gpurun_out/r02_readbw.txt 2>&1; cat gpurun_out/r02_readbw.txt
bash tools/gpu_k0_variants.sh
