#!/bin/bash
# round 2 validation, one GPU: full GPU suite, smoke, both bench arms, launch list of a step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','gpu_launches','clocks','roofline','roofline_eval_geometry','e2e_from_jpeg','eval_forward','swin_eval_forward','swin_train','vitti_configs','roofline_vit_step','host_decode','cpu_baseline','torch_b200_baseline'):
    print(k, json.dumps(d.get(k))[:600])
PY
tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo "ref exit $?"; cut -c1-600 gpurun_out/r02_bench_ref_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_step_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --no-swin > gpurun_out/r02_step_ncu.log 2>&1; echo "ncu exit $?"
python tools/launch_breakdown.py gpurun_out/r02_step_launches.csv gpurun_out/r02_step_launches.txt 2>&1 | head -24 | cut -c1-180
