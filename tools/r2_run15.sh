#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/dbg_diff.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_edge_list.py tests/test_gpu_baseline_configs.py -m gpu -q --no-header -p no:cacheprovider --timeout 300 2>&1 | tail -3
SKEWS=0 bash tools/r2_quick.sh 2>&1 | tail -2
ONLY=c4 DTYPES=bf16 NOREF=1 python tools/bench_configs.py 2>&1 | tail -1
EGNN_B200_SKEW_NS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair -s 3 -c 1 -f -o gpurun_out/prof_tc_pair_v7 \
    python bench.py --dtype bf16 --steps 2 --warmup 3 --lean > gpurun_out/ncu_tc_pair_v7.log 2>&1
