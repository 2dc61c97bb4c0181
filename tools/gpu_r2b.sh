#!/bin/bash
# round 2, call B: K0 v2 parity + timing + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_k0_gpu.py tests/test_vit_gpu.py::test_config0_jpeg_to_logits_end_to_end -m gpu -q --tb=short > gpurun_out/r02_pytest_k0v2.log 2>&1; echo "pytest k0 exit $?" | tee -a gpurun_out/r02_pytest_k0v2.log
tail -15 gpurun_out/r02_pytest_k0v2.log
timeout 300 python tools/k0_prof.py 20 > gpurun_out/r02_k0v2_time.json 2> gpurun_out/r02_k0v2_time.err; cat gpurun_out/r02_k0v2_time.json; tail -3 gpurun_out/r02_k0v2_time.err
RGBNM_K0_V1=1 timeout 300 python tools/k0_prof.py 20 > gpurun_out/r02_k0v1_time.json 2>/dev/null; cat gpurun_out/r02_k0v1_time.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k0_vit2 -s 1 -c 1 -o gpurun_out/r02_k0v2_eval -f python tools/k0_prof.py 1 > gpurun_out/r02_ncu_eval.log 2>&1; tail -2 gpurun_out/r02_ncu_eval.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k0_vit2 -s 21 -c 1 -o gpurun_out/r02_k0v2_train -f python tools/k0_prof.py 1 > gpurun_out/r02_ncu_train.log 2>&1; tail -2 gpurun_out/r02_ncu_train.log
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_compat_launcher_gpu.py -m gpu -q --tb=long > gpurun_out/r02_pytest_b.log 2>&1; tail -5 gpurun_out/r02_pytest_b.log
