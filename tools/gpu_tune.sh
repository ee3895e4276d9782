#!/bin/bash
# correctness first, then the tuning knobs of the fixed-cost stages at 2^20 (exact SRS = the bench's shape) and end to end
tag=${1:-tune}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_sort.py -x -q -m gpu > $o/${tag}_sort.log 2>&1; echo "sort rc=$?"; tail -2 $o/${tag}_sort.log
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_pst13.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
timeout 600 python -m pytest tests/test_gpu_rep3.py tests/test_gpu_field_curve.py -x -q -m gpu > $o/${tag}_rep3.log 2>&1; echo "rep3 rc=$?"; tail -3 $o/${tag}_rep3.log
timeout 300 python tools/bench_rep3.py > $o/${tag}_bench_rep3.log 2>&1; echo "bench_rep3 rc=$?"; tail -12 $o/${tag}_bench_rep3.log
S="python tools/sweep.py --exact --sizes 16,20,22 --dists uniform --steps 5"
echo "== base";            timeout 300 $S 2>&1 | grep "2^" | tee $o/${tag}_base.log
for u in 8 16; do echo "== acc_chunk_up $u"; timeout 300 $S --acc-chunk-up $u 2>&1 | grep "2^"; done | tee $o/${tag}_up.log
for g in 2 4 16; do echo "== group_l $g"; timeout 300 $S --group-l $g 2>&1 | grep "2^"; done | tee $o/${tag}_gl.log
H="python tools/sweep.py --exact --host --sizes 20,22 --dists uniform --steps 5"
echo "== e2e default";     timeout 300 $H 2>&1 | grep "2^" | tee $o/${tag}_e2e.log
for c in 2 3 4; do echo "== e2e stream from 2^20, chunks $c"; timeout 300 $H --stream-min 1048576 --stream-chunks $c 2>&1 | grep "2^"; done | tee -a $o/${tag}_e2e.log
