#!/bin/bash
# multi-GPU bench on this box: usage tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; tail -4 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
