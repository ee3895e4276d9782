#!/bin/bash
tag=${1:-first}
o=gpurun_out
mkdir -p $o
for pct in 30 50 67 80 100; do
  echo "=== 2 chunks, first chunk ${pct} % of an equal share"
  python tools/sweep.py --exact --host --sizes 20,22,24 --dists uniform --steps 7 --stream-chunks 2 --stream-first-pct $pct 2>&1 | grep "2^"
done | tee $o/${tag}_first.log
for pct in 50 75 100; do
  echo "=== 3 chunks, first chunk ${pct} % of an equal share"
  python tools/sweep.py --exact --host --sizes 20,22,24 --dists uniform --steps 7 --stream-chunks 3 --stream-first-pct $pct 2>&1 | grep "2^"
done | tee -a $o/${tag}_first.log
