#!/bin/bash
# round 2: full GPU suite + smoke + both bench arms at N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -15 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"; tail -c 6000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo "ref exit $?"; cat gpurun_out/r02_bench_ref_n1.json; tail -3 gpurun_out/r02_bench_ref_n1.err
