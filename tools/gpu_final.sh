#!/bin/bash
# round-end style check: full GPU test-suite, smoke, bench (our arm with cpu baseline), reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --arch vitti --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_vitti.json 2> gpurun_out/bench_vitti.err; echo "bench vitti exit $?"; tail -3 gpurun_out/bench_vitti.err; cut -c1-330 gpurun_out/bench_vitti.json
