#!/bin/bash
# round 2: K0 with the dynamic quad queue + 1024-thread statistics pre-pass: parity, timing, ncu, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_k0_gpu.py tests/test_k0_swin_gpu.py tests/test_vit_gpu.py::test_config0_jpeg_to_logits_end_to_end -m gpu -q --tb=short > gpurun_out/r02_pytest_k0dyn.log 2>&1; echo "pytest k0 exit $?" | tee -a gpurun_out/r02_pytest_k0dyn.log
tail -5 gpurun_out/r02_pytest_k0dyn.log
timeout 300 python tools/k0_prof.py 30 > gpurun_out/r02_k0dyn_time.json 2> gpurun_out/r02_k0dyn_time.err; cat gpurun_out/r02_k0dyn_time.json; tail -3 gpurun_out/r02_k0dyn_time.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k0_vit2 -s 1 -c 1 -o gpurun_out/r02_k0dyn_eval -f python tools/k0_prof.py 1 > gpurun_out/r02_ncu_eval.log 2>&1; tail -2 gpurun_out/r02_ncu_eval.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k0_vit2 -s 21 -c 1 -o gpurun_out/r02_k0dyn_train -f python tools/k0_prof.py 1 > gpurun_out/r02_ncu_train.log 2>&1; tail -2 gpurun_out/r02_ncu_train.log
timeout 600 ncu --set full --clock-control none -k regex:dcstats -s 2 -c 1 -o gpurun_out/r02_dcstats_train -f python tools/k0_prof.py 1 > gpurun_out/r02_ncu_dcstats.log 2>&1; tail -2 gpurun_out/r02_ncu_dcstats.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_b.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','roofline_eval_geometry','e2e_from_jpeg','swin_train','vitti_configs'):
    print(k, json.dumps(d.get(k))[:700])
PY
tail -3 gpurun_out/r02_bench_n1_b.err
