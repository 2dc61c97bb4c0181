#!/bin/bash
# A/B of library builds inside one box on the Swin eval forward: default vs build_alt/*.so (RGBNM_LIB override)
mkdir -p gpurun_out
for i in 1 2; do
  timeout 200 python tools/swin_bench.py --steps 10 2>/dev/null | grep -o '"ms_per_batch": [0-9.]*' | sed "s/^/default /"
  for f in build_alt/*.so; do
    RGBNM_LIB=$PWD/$f timeout 200 python tools/swin_bench.py --steps 10 2>/dev/null | grep -o '"ms_per_batch": [0-9.]*' | sed "s|^|$f |"
  done
done | tee gpurun_out/ab_swin.log
