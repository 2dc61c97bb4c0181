#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_k0_swin_gpu.py tests/test_k0_gpu.py -m gpu -q --tb=short > gpurun_out/pytest_k0_swin.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_k0_swin.log
tail -25 gpurun_out/pytest_k0_swin.log
timeout 300 python tools/swin_gemm_probe.py 8 > gpurun_out/swin_gemm_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/swin_gemm_probe.log | tail -30
timeout 300 python tools/k0_swin_time.py > gpurun_out/k0_swin_time.json 2> gpurun_out/k0_swin_time.err; echo "time exit $?"; cat gpurun_out/k0_swin_time.json; tail -3 gpurun_out/k0_swin_time.err
