#!/bin/bash
# round 2, GPU call 1: pipe rates, state check, baseline profile captures (with source) of the three tensor-core kernels
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep "Model name" >> gpurun_out/nproc.txt
./tools/pipe_bench > gpurun_out/pipe_bench.txt 2>&1; cat gpurun_out/pipe_bench.txt
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --dtype bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 1500 gpurun_out/bench_bf16.json; tail -3 gpurun_out/bench_bf16.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair -s 3 -c 1 -f -o gpurun_out/prof_tc_pair \
    python bench.py --dtype bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc_pair.log 2>&1
ONLY=c4 DTYPES=bf16 NOREF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_knn -s 2 -c 1 -f -o gpurun_out/prof_tc_knn \
    python tools/bench_configs.py > gpurun_out/ncu_tc_knn.log 2>&1
ONLY=c2 DTYPES=bf16 NOREF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 8 -c 4 -f -o gpurun_out/prof_tc_gemm \
    python tools/bench_configs.py > gpurun_out/ncu_tc_gemm.log 2>&1
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
ls -la gpurun_out
