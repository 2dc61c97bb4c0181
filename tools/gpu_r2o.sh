#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_swin_bwd_gpu.py -m gpu -q --tb=short -k "flat_engine or stage_learns" 2>&1 | tail -25
timeout 300 python tools/swin_train_bench.py --stage --steps 20 2>&1 | tail -3
timeout 300 python tools/swin_train_bench.py --stage --no-graph --steps 10 2>&1 | tail -1
