#!/bin/bash
# One gpurun payload: bring-up sweep, GPU tests, sanitizer pass, smoke, benches.  Logs land in gpurun_out/.
# Env knobs: BRINGUP=1 SANITIZE=1 EAGER=1 BENCH=1
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep "Model name" >> gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
if [ "${BRINGUP:-0}" = "1" ]; then
timeout 1200 python tools/tc_bringup.py > gpurun_out/bringup.log 2>&1; cat gpurun_out/bringup.log
fi
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 600 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
if [ "${SANITIZE:-0}" = "1" ]; then
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast.py -m gpu -q --no-header -p no:cacheprovider \
   -k "dense_everything or knn_k33 or net_c5_xavier or knn_select or adj_expand or dense_mask_padded or d64_n160 or d72_clamp or gemm_standalone" > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer.log
tail -3 gpurun_out/sanitizer.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_grad.py -m gpu -q --no-header -p no:cacheprovider \
   -k "oracle_fp64 and (dense_everything or knn_edges_mask or net_c5_xavier or adj_sparse_random or dense_fourier or knn_k_eq_n) or fp32 and dense_mdim32" > gpurun_out/sanitizer_grad.log 2>&1
echo "sanitizer(grad) exit $?" >> gpurun_out/sanitizer_grad.log
tail -3 gpurun_out/sanitizer_grad.log
fi
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -6 gpurun_out/smoke.log
if [ "${BENCH:-1}" = "1" ]; then
timeout 600 python bench.py --dtype fp32 --steps 5 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 2500 gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
timeout 600 python bench.py --dtype bf16 --steps 20 --warmup 5 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 2500 gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err
fi
if [ "${TRAIN:-0}" = "1" ]; then
(timeout 200 python tools/train_bench.py --iters 5; timeout 200 python tools/train_bench.py --b 8 --n 4096 --dim 256 --k 32 --edge-dim 4 --ref-b 1 --iters 5; timeout 100 python tools/train_bench.py --b 1 --n 16 --iters 20) 2>&1 | grep arm | tee gpurun_out/train_bench.jsonl
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py --iters 1 --no-ref > /dev/null 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_knn.csv python tools/train_bench.py --b 8 --n 4096 --dim 256 --k 32 --edge-dim 4 --iters 1 --no-ref > /dev/null 2>&1
fi
if [ "${EAGER:-0}" = "1" ]; then
./tools/pipe_bench > gpurun_out/pipe_bench.txt 2>&1; cat gpurun_out/pipe_bench.txt
timeout 600 python bench.py --impl eager --dtype bf16 --steps 5 --warmup 2 > gpurun_out/eager_bf16.json 2> gpurun_out/eager_bf16.err; cat gpurun_out/eager_bf16.json; tail -3 gpurun_out/eager_bf16.err
timeout 600 python bench.py --impl eager --dtype fp32 --steps 3 --warmup 1 > gpurun_out/eager_fp32.json 2> gpurun_out/eager_fp32.err; cat gpurun_out/eager_fp32.json; tail -3 gpurun_out/eager_fp32.err
fi
if [ "${CONFIGS:-0}" = "1" ]; then
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
fi
