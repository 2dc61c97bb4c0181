#!/bin/bash
# Round-end validation on the GPU box: every GPU test, smoke, the Swin eval bench, the headline bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -4 gpurun_out/smoke.log
timeout 300 python tools/swin_bench.py --steps 10 > gpurun_out/swin_bench.json 2> gpurun_out/swin_bench.err; echo "swin bench exit $?"; cat gpurun_out/swin_bench.json; tail -3 gpurun_out/swin_bench.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cut -c1-3000 gpurun_out/bench.json
