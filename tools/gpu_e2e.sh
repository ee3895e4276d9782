#!/bin/bash
# chunk-count sweep of the chunked pipeline: pinned host scalars (end to end) and device-resident scalars
tag=${1:-e2e}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
for c in 2 3 4 6 8; do
  echo "=== host chunks $c" >> $o/${tag}_e2e.log
  python tools/sweep.py --exact --host --sizes 20,22,24 --dists uniform --steps 5 --stream-chunks $c 2>&1 | grep "2^" >> $o/${tag}_e2e.log
  echo "=== resident chunks $c" >> $o/${tag}_e2e.log
  python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 5 --stream-chunks $c 2>&1 | grep "2^" >> $o/${tag}_e2e.log
done
echo "=== resident one-shot" >> $o/${tag}_e2e.log
python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 5 --stream-chunks 1 2>&1 | grep "2^" >> $o/${tag}_e2e.log
cat $o/${tag}_e2e.log
