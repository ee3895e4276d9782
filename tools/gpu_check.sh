#!/bin/bash
# One gpurun call: the sort tests first (fail fast), the whole gpu tier, smoke, one bench line.  Output under gpurun_out/<tag>_*.
tag=${1:-chk}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_sort.py -x -q -m gpu > $o/${tag}_sort.log 2>&1; echo "sort rc=$?" | tee -a $o/${tag}_sort.log
tail -5 $o/${tag}_sort.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_sort.py --durations=15 > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?" | tee -a $o/${tag}_gpu.log
tail -30 $o/${tag}_gpu.log
timeout 300 python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $o/${tag}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"; tail -c 3000 $o/${tag}_bench.json; tail -5 $o/${tag}_bench.err
