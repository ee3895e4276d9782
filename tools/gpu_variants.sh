#!/bin/bash
# level-1 accumulate kernel built with other block sizes / register budgets (libcozk_msm_acc<block>_<maxreg>.so), side by side
tag=${1:-var}
o=gpurun_out
mkdir -p $o
for lib in "" $(ls co-zkvms_b200/libcozk_msm_acc*.so 2>/dev/null); do
  if [ -n "$lib" ]; then export COZK_LIB=$PWD/$lib; v=$(basename $lib .so); else unset COZK_LIB; v=default; fi
  echo "=== $v"
  timeout 300 python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 5 2>&1 | grep "2^"
done | tee $o/${tag}_variants.log
unset COZK_LIB
