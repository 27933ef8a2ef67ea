#!/bin/bash
set -u
mkdir -p gpurun_out
SKEWS="0" bash tools/r2_quick.sh
EGNN_B200_SKEW_NS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair -s 3 -c 1 -f -o gpurun_out/prof_tc_pair_1wg \
    python tools/one_wg.py 128 > gpurun_out/ncu_1wg.log 2>&1
ls -la gpurun_out/prof_tc_pair_1wg.ncu-rep
