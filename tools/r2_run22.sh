#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider --timeout 300 2>&1 | tail -3
timeout 300 python tools/rowrange_time.py
SKEWS=0 bash tools/r2_quick.sh 2>&1 | tail -1
