#!/bin/bash
tag=${1:-open}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_rep3.py tests/test_gpu_pst13.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
for nv in 16 20 22; do
  for r in 0 1; do
    echo "=== nv $nv open_small_ragged $r"
    COZK_OPEN_TRACE=1 timeout 600 python tools/bench_rep3.py --log2n 16 --k 2 --nv $nv --small 15 --small-ragged $r 2>&1 | grep "\"open\"\|\[open\]" | tail -5 | cut -c1-330
  done
done | tee $o/${tag}_open.log
