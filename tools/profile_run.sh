#!/bin/bash
# One gpurun call that produces everything profiles/ summarises for the headline path (1 GPU):
#   bash tools/profile_run.sh <tag>      -> gpurun_out/<tag>_*
# bench line (+ reference arm), size sweep, then - each only after the same command exited 0 without ncu - the launch
# list and `--set full` captures of the accumulate kernel and of the sort kernels.
tag=${1:-prof}
o=gpurun_out
mkdir -p $o
python bench.py --steps 20 --warmup 5 > $o/${tag}_bench_1gpu.json 2> $o/${tag}_bench.err || { tail -5 $o/${tag}_bench.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_bench.err
python tools/sweep.py --exact --sizes 12,16,18,20,22,24,25,26 --dists uniform,const,wminus --steps 3 2>&1 | grep "2^" > $o/${tag}_sweep.log
python tools/sweep.py --exact --host --sizes 20,22,24,26 --dists uniform --steps 3 2>&1 | grep "2^" > $o/${tag}_sweep_e2e.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-log2n 0 --replay-log2t 0"
$B > $o/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches.csv $B > $o/${tag}_ncu1.log 2>&1
$B > $o/${tag}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 10 -c 2 -o $o/${tag}_prof_accumulate -f $B > $o/${tag}_ncu2.log 2>&1
$B > $o/${tag}_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_sort -s 40 -c 5 -o $o/${tag}_prof_sort -f $B > $o/${tag}_ncu3.log 2>&1
ls -la $o | tail -14
