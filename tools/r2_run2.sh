#!/bin/bash
# round 2, GPU call 2: persistent tc_pair kernel -- parity, skew sweep, mix variants
set -u
mkdir -p gpurun_out
./tools/pipe_bench 2>&1 | grep -E "round mix" > gpurun_out/pipe_bench_mix.txt; cat gpurun_out/pipe_bench_mix.txt
timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 -x -s > gpurun_out/pytest_fast.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fast.log
grep -E "err|passed|failed|exit|Error|assert" gpurun_out/pytest_fast.log | tail -60
for SK in 0 8000 16000 24000 32000 48000; do
  EGNN_B200_SKEW_NS=$SK timeout 300 python bench.py --dtype bf16 --steps 20 --warmup 5 --lean > gpurun_out/bench_skew_$SK.json 2> gpurun_out/bench_skew_$SK.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_skew_$SK.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('skew $SK', 'ms/step %.4f'%d['ms_per_step'], 'edge %.4f'%r['launch_ms'], 'frac %.4f'%r['frac'], r['stage_ms_per_step'])
except Exception as e:
    print('skew $SK failed', e, open('gpurun_out/bench_skew_$SK.err').read()[-600:])
PY
done
