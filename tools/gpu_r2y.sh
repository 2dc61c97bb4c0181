#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gemm_check.py all > gpurun_out/r02_gemm_check.log 2>&1; echo "gemm_check exit $?"
grep -E "FAIL|^time (fc1|dfc1|dgelu|gelu|fc2 )|GELU" gpurun_out/r02_gemm_check.log | head -30
timeout 600 python -m pytest tests/test_vit_gpu.py tests/test_swin_gpu.py -m gpu -q --tb=short 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu --no-swin 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
