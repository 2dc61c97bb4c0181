#!/bin/bash
# round 2, two GPUs: NCCL replica test + torchrun bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench n2 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n4.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','e2e','e2e_from_jpeg','swin_train','roofline'):
    print(k, json.dumps(d.get(k))[:400])
PY
tail -3 gpurun_out/r02_bench_n4.err
