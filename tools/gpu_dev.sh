#!/bin/bash
# Development driver run on the GPU box through gpurun: tests + smoke + bench (+ ncu).
# usage: tools/gpu_dev.sh [tests|notests] [ncu-kernel-regex] [pytest -k expr]
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt
if [ "${1:-tests}" = "tests" ]; then
timeout 1200 python -m pytest tests -m gpu -q --tb=short ${3:+-k "$3"} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_bench.log 2>&1
if [ -n "$2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 6 -c 2 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
