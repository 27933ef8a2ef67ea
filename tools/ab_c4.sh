#!/bin/bash
# same-box A/B of tc_knn variants on BASELINE config 4
set -u
for rep in 1 2; do
  for v in egnn_pytorch_b200/lib/variants/*.so; do
    echo -n "rep $rep $(basename $v): "; EGNN_B200_LIB=$PWD/$v timeout 300 python tools/c4_stages.py 2>&1 | tail -1
  done
done
