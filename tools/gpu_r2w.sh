#!/bin/bash
# round 2: compute-sanitizer memcheck / racecheck over the kernels written this round
mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_k0_gpu.py -m gpu -q -x -k "golden_pipeline or eval_geometry or without_subblock or freq_enhance or properties_full_batch" > gpurun_out/r02_sanitize_k0.log 2>&1; echo "memcheck k0 exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/r02_sanitize_k0.log | head -8
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_swin_bwd_gpu.py -m gpu -q -x -k "window_attention_bwd or layernorm_res or stage_learns" > gpurun_out/r02_sanitize_swin.log 2>&1; echo "memcheck swin exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/r02_sanitize_swin.log | head -8
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_vit_gpu.py -m gpu -q -x -k "embedding_variants" > gpurun_out/r02_sanitize_vit.log 2>&1; echo "memcheck vit exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/r02_sanitize_vit.log | head -8
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_k0_gpu.py tests/test_swin_bwd_gpu.py -m gpu -q -x -k "eval_geometry or (window_attention_bwd and 16-3-4) or (layernorm_res and 96)" > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/r02_racecheck.log | head -12
