#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_k0_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_k0.log 2>&1; echo "pytest k0 exit $?" | tee -a gpurun_out/pytest_k0.log
tail -3 gpurun_out/pytest_k0.log
timeout 300 python bench.py --stage k0 --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_k0.json 2> gpurun_out/bench_k0.err; echo "bench k0 exit $?"; tail -3 gpurun_out/bench_k0.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k0.json'))
print("train mix:", d["roofline"]["kernel_ms"], "ms frac", d["roofline"]["frac"]); print("eval:", d["roofline_eval_geometry"]["kernel_ms"], "ms frac", d["roofline_eval_geometry"]["frac"])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cut -c1-1800 gpurun_out/bench.json
