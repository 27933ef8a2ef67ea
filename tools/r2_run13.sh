#!/bin/bash
set -u
mkdir -p gpurun_out
echo skip tests
time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 6000 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_ref.json
