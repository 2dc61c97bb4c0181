#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_feeder_gpu.py -m gpu -q --tb=short > gpurun_out/pytest_feeder.log 2>&1; echo "pytest feeder exit $?"; tail -15 gpurun_out/pytest_feeder.log
timeout 300 python examples/train_synthetic_dct.py --steps 60 --batch 64 > gpurun_out/example.log 2>&1; echo "example exit $?"; tail -8 gpurun_out/example.log
