#!/bin/bash
tag=${1:-reg}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
timeout 600 python tools/bench_register.py --sizes 12,16,20,22,24 2>&1 | tee $o/${tag}_register.log
