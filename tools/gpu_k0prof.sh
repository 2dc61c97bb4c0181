#!/bin/bash
mkdir -p gpurun_out
# launches: warmup(3)+steps(2) train-mix device loop, same e2e, then eval loop: capture one train-mix (skip 2) and the eval ones (skip ~ 2*(3+2)+3)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k0_fused_kernel -s 2 -c 1 -f -o gpurun_out/prof_k0_train python bench.py --stage k0 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_k0a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k0_fused_kernel -s 14 -c 1 -f -o gpurun_out/prof_k0_eval python bench.py --stage k0 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_k0b.log 2>&1
tail -2 gpurun_out/ncu_k0b.log
