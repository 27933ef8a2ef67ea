#!/bin/bash
# round 2 evidence collection on one B200: tests, smoke, bench line, launch list, ncu --set full of the three tensor-core kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json; tail -2 gpurun_out/bench_full.err
timeout 300 python tools/tiny_stages.py > gpurun_out/tiny_stages.txt 2>&1; cat gpurun_out/tiny_stages.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bf16.csv \
    python bench.py --dtype bf16 --steps 2 --warmup 3 --lean > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair -s 3 -c 1 -f -o gpurun_out/prof_tc_pair_final \
    python bench.py --dtype bf16 --steps 2 --warmup 3 --lean > gpurun_out/ncu_tc_pair.log 2>&1
ONLY=c4 DTYPES=bf16 NOREF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_knn -s 2 -c 1 -f -o gpurun_out/prof_tc_knn_final \
    python tools/bench_configs.py > gpurun_out/ncu_tc_knn.log 2>&1
ONLY=c2 DTYPES=bf16 NOREF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 6 -c 3 -f -o gpurun_out/prof_tc_gemm_final \
    python tools/bench_configs.py > gpurun_out/ncu_tc_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
