#!/bin/bash
# round 2, call A: new parity tests + bench line with the reference-itself arms
mkdir -p gpurun_out
python -m pytest tests/test_vit_gpu.py tests/test_ddp_gpu.py tests/test_compat_launcher_gpu.py tests/test_k0_gpu.py tests/test_feeder_gpu.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r02_pytest_a.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -5 gpurun_out/r02_pytest_a.log
tail -c 1500 gpurun_out/r02_bench_a.json
