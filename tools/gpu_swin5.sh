#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -q --tb=short > gpurun_out/pytest_swin.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_swin.log
tail -30 gpurun_out/pytest_swin.log
timeout 300 python tools/swin_bench.py --steps 10 > gpurun_out/swin_bench.json 2> gpurun_out/swin_bench.err; echo "bench exit $?"; cat gpurun_out/swin_bench.json; tail -5 gpurun_out/swin_bench.err
RGBNM_SWIN_FUSE_MAX_DIM=384 timeout 300 python tools/swin_bench.py --steps 10 > gpurun_out/swin_bench_fuse384.json 2>/dev/null; cut -c1-200 gpurun_out/swin_bench_fuse384.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/swin_launches.csv python tools/swin_bench.py --steps 1 --no-graph > gpurun_out/swin_ncu.log 2>&1; echo "ncu exit $?"
python tools/launch_breakdown.py gpurun_out/swin_launches.csv 2>&1 | head -24
