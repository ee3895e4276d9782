#!/bin/bash
# batched-affine pre-reduction: parity tests, the rounds on their own, and the MSM with 0 / 2 / 3 / 4 rounds
tag=${1:-aff}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_affine.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -5 $o/${tag}_gpu.log
timeout 600 python tools/bench_affine.py --log2n 20 > $o/${tag}_rounds.log 2>&1; cat $o/${tag}_rounds.log | tail -6
for r in 0 2 3 4; do
  echo "=== affine rounds $r" >> $o/${tag}_sweep.log
  timeout 600 python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 5 --affine-rounds $r 2>&1 | grep "2^" >> $o/${tag}_sweep.log
done
cat $o/${tag}_sweep.log
