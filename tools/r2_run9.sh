#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fast.py -m gpu -q --no-header -p no:cacheprovider --timeout 300 2>&1 | tail -3
for CFG in c1 c3 c5; do
  for DT in fp32 bf16; do
    ONLY=$CFG DTYPES=$DT NOREF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${CFG}_${DT}.csv \
      python tools/bench_configs.py > gpurun_out/ncu_${CFG}_${DT}.log 2>&1
  done
done
ls gpurun_out/launches_c*
