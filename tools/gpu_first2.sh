#!/bin/bash
tag=${1:-first2}
o=gpurun_out
mkdir -p $o
for pct in 140 160 175 185 193; do
  echo "=== resident, 2 chunks, first chunk ${pct} % of an equal share"
  python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 7 --chunk-min 1048576 --stream-chunks 2 --stream-first-pct $pct 2>&1 | grep "2^"
done | tee $o/${tag}_first.log
echo "=== resident one piece" | tee -a $o/${tag}_first.log
python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 7 2>&1 | grep "2^" | tee -a $o/${tag}_first.log
