#!/bin/bash
tag=${1:-tune2}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_rep3.py tests/test_gpu_field_curve.py -x -q -m gpu > $o/${tag}_rep3.log 2>&1; echo "rep3 rc=$?"; tail -3 $o/${tag}_rep3.log
timeout 300 python tools/bench_rep3.py > $o/${tag}_bench_rep3.log 2>&1; echo "bench_rep3 rc=$?"; tail -12 $o/${tag}_bench_rep3.log | cut -c1-600
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_pst13.py tests/test_gpu_sort.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
S="python tools/sweep.py --exact --sizes 16,18,20,22,24 --dists uniform --steps 5"
echo "== base";            timeout 300 $S 2>&1 | grep "2^" | tee $o/${tag}_base.log
for g in 32 64; do echo "== group_l $g"; timeout 300 python tools/sweep.py --exact --sizes 22,24 --dists uniform --steps 3 --group-l $g 2>&1 | grep "2^"; done | tee $o/${tag}_gl.log
echo "== 2^18 group_l 4/16"; for g in 4 16; do timeout 300 python tools/sweep.py --exact --sizes 18 --dists uniform --steps 5 --group-l $g 2>&1 | grep "2^"; done | tee -a $o/${tag}_gl.log
echo "== e2e"; timeout 300 python tools/sweep.py --exact --host --sizes 20,22,24 --dists uniform --steps 3 2>&1 | grep "2^" | tee $o/${tag}_e2e.log
