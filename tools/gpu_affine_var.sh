#!/bin/bash
# batched-affine kernel variants (libcozk_msm_aff<K>_<minblocks>[p].so), one round in front of the accumulate levels
tag=${1:-affv}
o=gpurun_out
mkdir -p $o
for lib in "" $(ls co-zkvms_b200/libcozk_msm_aff*.so 2>/dev/null); do
  if [ -n "$lib" ]; then export COZK_LIB=$PWD/$lib; v=$(basename $lib .so); else unset COZK_LIB; v=default; fi
  echo "=== $v"
  timeout 300 python tools/sweep.py --exact --sizes 20,22 --dists uniform --steps 5 --affine-rounds 1 2>&1 | grep "2^"
done | tee $o/${tag}_variants.log
unset COZK_LIB
