#!/bin/bash
# tests + bench + ncu launch list of one eager step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
