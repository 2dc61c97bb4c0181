#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench PDL on exit $?"; tail -3 gpurun_out/bench.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench.json
RGBNM_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench PDL off exit $?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_nopdl.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench PDL on (2nd) exit $?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench2.json
