#!/bin/bash
# A/B of VAR=VALUE inside one box: bench step time (2 repetitions) + GEMM timings.  usage: tools/gpu_ab_env.sh VAR=VALUE
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/default /"
  env "$1" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/$1 /"
done
env "$1" timeout 300 python tools/gemm_check.py big 2>&1 | grep -E "FAIL|^time [a-z0-9]+:" | sed "s/^/$1 /"
