#!/bin/bash
tag=${1:-r4}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_rep3.py -x -q -m gpu > $o/${tag}_rep3.log 2>&1; echo "rep3 rc=$?"; tail -2 $o/${tag}_rep3.log
for w in 1 2 4; do for lg in 20 22; do
  echo "== chi_waves=$w log2n=$lg"; timeout 300 python tools/bench_rep3.py --chi-waves $w --log2n $lg --k 32 --nv 16 2>/dev/null | grep -E '"experiment": "(ingest|chi)"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ('experiment', 'ingest_kernel_ms', 'ingest_frac_of_hbm', 'ms', 'frac_of_hbm', 'frac_of_imad', 'gbs')})"
done; done | tee $o/${tag}_chi.log
for lib in "" libcozk_msm_v512_8.so libcozk_msm_v256_32.so; do
  if [ -n "$lib" ]; then export COZK_LIB=$PWD/co-zkvms_b200/$lib; v=$lib; else unset COZK_LIB; v=default; fi
  echo "=== $v"
  timeout 300 python tools/sweep.py --exact --sizes 16,20,22,24 --dists uniform --steps 5 2>&1 | grep "2^" | tee $o/${tag}_sweep_$v.log
done
unset COZK_LIB
