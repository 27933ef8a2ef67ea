#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 -s > gpurun_out/pytest_fast.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fast.log
grep -E "err|passed|failed|exit|Error|assert" gpurun_out/pytest_fast.log | tail -70
echo "== saturation skew=0"; EGNN_B200_SKEW_NS=0 timeout 300 python tools/wg_saturation.py 2>&1 | tail -9
echo "== saturation skew=0 no coors"; EGNN_B200_SKEW_NS=0 timeout 300 python tools/wg_saturation.py --no-coors 2>&1 | tail -9
echo "== saturation skew=24000"; EGNN_B200_SKEW_NS=24000 timeout 300 python tools/wg_saturation.py 2>&1 | tail -9
for SK in 0 24000; do
EGNN_B200_SKEW_NS=$SK timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair -s 3 -c 1 -f -o gpurun_out/prof_tc_pair_v6_skew$SK \
    python bench.py --dtype bf16 --steps 2 --warmup 3 --lean > gpurun_out/ncu_tc_pair_v6_$SK.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
