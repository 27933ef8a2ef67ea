#!/bin/bash
set -u
mkdir -p gpurun_out
N=${NG:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc $?"; tail -c 2500 gpurun_out/bench_${N}gpu.json; tail -4 gpurun_out/bench_${N}gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | tail -c 300
