#!/bin/bash
mkdir -p gpurun_out
RGBNM_LIB=$PWD/build_alt/librgbnm_async33.so timeout 600 python -m pytest tests/test_k0_gpu.py -x -q 2>&1 | tail -3
bash tools/gpu_k0_variants.sh
