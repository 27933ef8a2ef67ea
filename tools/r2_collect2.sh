#!/bin/bash
# round 2, final evidence collection on one B200 (tc_pair / tc_gemm are unchanged since tools/r2_collect.sh ran: their ncu
# captures are kept; the neighbour-list kernels changed and are re-captured)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 300 gpurun_out/bench_full.json; tail -2 gpurun_out/bench_full.err
timeout 300 python tools/tiny_stages.py > gpurun_out/tiny_stages.txt 2>&1; cat gpurun_out/tiny_stages.txt
timeout 300 python tools/c4_stages.py > gpurun_out/c4_stages.txt 2>&1; cat gpurun_out/c4_stages.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bf16.csv \
    python bench.py --dtype bf16 --steps 2 --warmup 3 --lean > gpurun_out/ncu_launches.log 2>&1
ONLY=c4 DTYPES=bf16 NOREF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_knn|knn_warp_select" -s 4 -c 2 -f \
    -o gpurun_out/prof_c4_final python tools/bench_configs.py > gpurun_out/ncu_c4_final.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
