#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_vit_gpu.py -m gpu -q -x -k "attention_fwd_bwd and 2-3 or gemm_epilogues and 128 or wgrad_atomic and 1024 or layernorm and 384 or mixup" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/sanitize_memcheck.log | head -20
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_k0_gpu.py -m gpu -q -x -k "golden_pipeline or eval_geometry" > gpurun_out/sanitize_memcheck_k0.log 2>&1; echo "memcheck k0 exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/sanitize_memcheck_k0.log | head -10
