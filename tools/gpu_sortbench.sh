#!/bin/bash
# sort tests + bench lines for the default build and the variant builds of the sort (COZK_LIB), + launch list with ncu
tag=${1:-sb}
o=gpurun_out
mkdir -p $o
for lib in "" libcozk_msm_t256.so; do
  if [ -n "$lib" ]; then export COZK_LIB=$PWD/co-zkvms_b200/$lib; v=$lib; else unset COZK_LIB; v=default; fi
  [ -n "$lib" ] && [ ! -f "$COZK_LIB" ] && continue
  echo "=== $v"
  timeout 600 python -m pytest tests/test_gpu_sort.py -x -q -m gpu > $o/${tag}_sort_$v.log 2>&1; echo "sort rc=$?"; tail -2 $o/${tag}_sort_$v.log
  timeout 600 python bench.py --steps 10 --warmup 3 --strong-log2n 0 --no-cpu-baseline > $o/${tag}_bench_$v.json 2> $o/${tag}_bench_$v.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('$o/${tag}_bench_$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'])"
  timeout 600 python tools/sweep.py --sizes 16,20,22,24 --dists uniform,const,wminus --steps 3 2>&1 | grep "2^" | tee $o/${tag}_sweep_$v.log
done
unset COZK_LIB
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_pst13.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-log2n 0 > $o/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
