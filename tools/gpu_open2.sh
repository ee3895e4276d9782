#!/bin/bash
tag=${1:-open2}
o=gpurun_out
mkdir -p $o
for w in 0 9 10 11 12 13 14 15; do
  echo "=== nv 16 open_small_window $w"
  COZK_OPEN_TRACE=1 timeout 600 python tools/bench_rep3.py --log2n 16 --k 2 --nv 16 --small 15 --small-window $w 2>&1 | grep "\"open\"\|\[open\]" | tail -3 | cut -c1-330
done | tee $o/${tag}_open.log
