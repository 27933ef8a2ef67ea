#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q --no-header -p no:cacheprovider --timeout 300 2>&1 | tail -2
echo "== saturation"; EGNN_B200_SKEW_NS=0 timeout 300 python tools/wg_saturation.py 2>&1 | tail -9
EGNN_B200_SKEW_NS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair -s 3 -c 1 -f -o gpurun_out/prof_tc_pair_1wg \
    python tools/one_wg.py 128 > gpurun_out/ncu_1wg.log 2>&1
