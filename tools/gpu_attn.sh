#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_check.py > gpurun_out/attn_check.log 2>&1; echo "attn_check exit $?" | tee -a gpurun_out/attn_check.log
grep -E "FAIL|time|FAILS|rror" gpurun_out/attn_check.log | head -40
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cut -c1-420 gpurun_out/bench.json
