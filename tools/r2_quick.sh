#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 > gpurun_out/pytest_fast.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fast.log
grep -E "passed|failed|exit|Error|^E  " gpurun_out/pytest_fast.log | tail -20
for SK in ${SKEWS:-0}; do
  EGNN_B200_SKEW_NS=$SK timeout 300 python bench.py --dtype bf16 --steps 20 --warmup 5 --lean > gpurun_out/bench_skew_$SK.json 2> gpurun_out/bench_skew_$SK.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_skew_$SK.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('skew $SK', 'ms/step %.4f'%d['ms_per_step'], 'edge %.4f'%r['launch_ms'], 'frac %.4f'%r['frac'], r['stage_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e:
    print('skew $SK failed', e, open('gpurun_out/bench_skew_$SK.err').read()[-600:])
PY
done
