#!/bin/bash
# quick loop for the tensor-core path: fast tests + bf16 bench (+ optional ncu)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 -s > gpurun_out/pytest_fast.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fast.log
grep -E "err|passed|failed|exit" gpurun_out/pytest_fast.log | tail -25
timeout 600 python bench.py --dtype bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_bf16.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','dtype','gpu_launches','equivariance_err']})
    print('e2e',d['e2e']); r=d['roofline']; print(r['stage_ms_per_step'], 'sfu frac', r['frac'], 'eff ref TF', r['effective_reference_tflops'])
except Exception as e:
    print('bench failed', e, open('gpurun_out/bench_bf16.err').read()[-800:])
PY
if [ "${NCU:-0}" = "1" ]; then bash tools/profile.sh; fi
