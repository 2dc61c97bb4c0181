#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_swin_bwd_gpu.py tests/test_swin_gpu.py -m gpu -q --tb=short 2>&1 | tail -8
timeout 300 python tools/swin_train_bench.py --stage --steps 20 2>&1 | tail -1
