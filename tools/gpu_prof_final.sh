#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 4 -f -o gpurun_out/prof_attn python tools/attn_prof.py > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -c 5 -f -o gpurun_out/prof_gemm python tools/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_attn.log; tail -1 gpurun_out/ncu_gemm.log
