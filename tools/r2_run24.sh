#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropout.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 -x 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout 600 2>&1 | tail -4
REPS=1 bash tools/ab_variants.sh 2>&1 | tail -3
