#!/bin/bash
tag=${1:-l2}
o=gpurun_out
mkdir -p $o
for g in 128 64 32; do
  export COZK_L2_FETCH=$g
  for r in 0 1 3; do
    echo "=== L2 fetch granularity $g, affine rounds $r"
    python tools/sweep.py --exact --sizes 20,22 --dists uniform --steps 5 --affine-rounds $r 2>&1 | grep "2^"
  done
done | tee $o/${tag}_l2.log
