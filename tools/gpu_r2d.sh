#!/bin/bash
tag=${1:-r2d}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "row_column" > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
for r in 0 1; do
  echo "=== reduce_2d $r"
  python tools/sweep.py --exact --sizes 12,16,18,20,22,24 --dists uniform --steps 5 --reduce-2d $r 2>&1 | grep "2^"
done | tee $o/${tag}_sweep.log
