#!/bin/bash
# 2 GPUs: multi-GPU tests (NCCL + peer-memory all-gather), fast-path tests, bench with row_sharded block
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 > gpurun_out/pytest_fast.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fast.log
grep -E "passed|failed|exit|Error|^E  |FAILED" gpurun_out/pytest_fast.log | tail -30
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 500 -x > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -30 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 3000 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
