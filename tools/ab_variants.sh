#!/bin/bash
# same-box A/B of kernel variants: EGNN_B200_LIB selects the shared library; variants interleaved, REPS repetitions;
# CHECK=1 also runs the tensor-core parity tests against every variant
set -u
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-3}); do
  for v in egnn_pytorch_b200/lib/variants/*.so; do
    EGNN_B200_LIB=$PWD/$v timeout 300 python bench.py --dtype bf16 --steps 20 --warmup 5 --lean 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('rep $rep %-18s ms/step %.4f edge %.4f pre %.4f post %.4f' % ('$(basename $v)', d['ms_per_step'], r['launch_ms'], r['stage_ms_per_step']['node_pre'], r['stage_ms_per_step']['node_post']))"
  done
done
if [ "${CHECK:-0}" = "1" ]; then
  for v in egnn_pytorch_b200/lib/variants/*.so; do
    echo -n "check $(basename $v): "; EGNN_B200_LIB=$PWD/$v timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_edge_list.py -m gpu -q --no-header -p no:cacheprovider --timeout 300 2>&1 | tail -1
  done
fi
