#!/bin/bash
# ncu captures (1 GPU): launch list of a short bench run + full-set capture of the fused edge kernel.
set -u
mkdir -p gpurun_out
DT=${DT:-bf16}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${DT}.csv \
    python bench.py --dtype $DT --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${DT}.log 2>&1
KERN=${KERN:-tc_pair}
ncu --set full --clock-control none --import-source on -k regex:$KERN -s 3 -c 1 -f -o gpurun_out/prof_${KERN} \
    python bench.py --dtype $DT --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${KERN}.log 2>&1
ls -la gpurun_out/*.ncu-rep
