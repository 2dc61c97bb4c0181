#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_swin_stage_launches.csv python tools/swin_train_bench.py --stage --no-graph --steps 1 > gpurun_out/r02_swin_stage_ncu.log 2>&1; echo "ncu exit $?"
python tools/launch_breakdown.py gpurun_out/r02_swin_stage_launches.csv gpurun_out/r02_swin_train_v3_launches.txt 2>&1 | head -75 | cut -c1-200
