#!/bin/bash
# A/B of two builds of the library inside one box: default vs build_alt/librgbnm_b200_alt.so
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/default /"
  RGBNM_LIB=$PWD/build_alt/librgbnm_b200_alt.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/alt     /"
done
RGBNM_LIB=$PWD/build_alt/librgbnm_b200_alt.so timeout 300 python tools/gemm_check.py big 2>&1 | grep -E "FAIL|^time [a-z0-9]+:" | sed "s/^/alt /"
timeout 300 python tools/gemm_check.py big 2>&1 | grep -E "FAIL|^time [a-z0-9]+:" | sed "s/^/def /"
