#!/bin/bash
tag=${1:-r7}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "streamed or sliced or multi_device or baseline_sizes or properties" > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
H="python tools/sweep.py --exact --host --dists uniform --steps 5"
echo "== e2e default";  timeout 300 $H --sizes 20,21,22,24 2>&1 | grep "2^" | tee $o/${tag}_e2e.log
for c in 2 3 4; do echo "== e2e stream from 2^20, chunks $c"; timeout 300 $H --sizes 20,21,22 --stream-min 1048576 --stream-chunks $c 2>&1 | grep "2^"; done | tee -a $o/${tag}_e2e.log
for c in 4 8; do echo "== e2e 2^24/2^26 chunks $c"; timeout 300 $H --sizes 24,26 --stream-chunks $c 2>&1 | grep "2^"; done | tee -a $o/${tag}_e2e.log
