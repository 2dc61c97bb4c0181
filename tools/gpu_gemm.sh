#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gemm_check.py all > gpurun_out/gemm_check_v2.log 2>&1; echo "gemm_check v2 exit $?" | tee -a gpurun_out/gemm_check_v2.log
grep -E "FAIL|^time|FAILS" gpurun_out/gemm_check_v2.log | head -60
RGBNM_GEMM_NO384=1 timeout 300 python tools/gemm_check.py big 2>&1 | grep -E "^time (fc2|dfc2|dqkv)" | sed 's/^/BN192: /'
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -c 5 -f -o gpurun_out/prof_gemm python tools/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
