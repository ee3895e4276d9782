#!/bin/bash
tag=${1:-open3}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_rep3.py tests/test_gpu_pst13.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
for nv in 16 18 20 22; do
  echo "=== nv $nv"
  COZK_OPEN_TRACE=1 timeout 600 python tools/bench_rep3.py --log2n 16 --k 2 --nv $nv --small 15 2>&1 | grep "experiment\": \"open\"\|\[open\]" | tail -4 | cut -c1-330
done | tee $o/${tag}_open.log
