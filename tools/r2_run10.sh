#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|exit|Error|^E  |FAILED" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8
timeout 300 python tools/tiny_stages.py 2>&1 | tail -8
