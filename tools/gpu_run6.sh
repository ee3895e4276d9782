#!/bin/bash
tag=${1:-r6}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_sort.py -x -q -m gpu > $o/${tag}_sort.log 2>&1; echo "sort rc=$?"; tail -3 $o/${tag}_sort.log
for b in 8 9 10 11; do echo "== sort_digit_bits $b"; timeout 300 python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 3 --sort-bits $b 2>&1 | grep "2^"; done | tee $o/${tag}_sortbits.log
