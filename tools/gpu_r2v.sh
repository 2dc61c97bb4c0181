#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_feeder_gpu.py tests/test_compat_launcher_gpu.py -m gpu -q --tb=short 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --no-swin > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_c.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_from_jpeg','host_decode','vitti_configs'):
    print(k, json.dumps(d.get(k))[:300])
PY
