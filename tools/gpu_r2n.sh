#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_swin_bwd_gpu.py -m gpu -q --tb=short 2>&1 | tail -15
timeout 300 python tools/swin_train_bench.py --steps 8 2>&1 | tail -2
RGBNM_WATTN_BWD_SIMT=1 timeout 300 python tools/swin_train_bench.py --steps 8 2>&1 | tail -1 | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_swin_train_launches2.csv python tools/swin_train_bench.py --steps 1 > gpurun_out/r02_swin_train_ncu2.log 2>&1; echo "ncu exit $?"
python tools/launch_breakdown.py gpurun_out/r02_swin_train_launches2.csv gpurun_out/r02_swin_train_v2_launches.txt 2>&1 | head -30 | cut -c1-210
