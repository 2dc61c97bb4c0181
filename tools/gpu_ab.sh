#!/bin/bash
# A/B of an environment switch inside one box: usage tools/gpu_ab.sh VAR
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/default   /"
  env $1=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/$1=1 /"
done
