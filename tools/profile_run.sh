#!/bin/bash
# One gpurun call that produces everything profiles/ summarises for the headline path (1 GPU):
#   bash tools/profile_run.sh <tag>      -> gpurun_out/<tag>_*
# bench line (+ reference arm), microbenchmarks, size sweep, then - each only after the same command exited 0 without
# ncu - the launch list and one `--set full` capture of the accumulate kernels.
tag=${1:-prof}
o=gpurun_out
python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_ref.json 2>> $o/${tag}_bench.err
python tools/microbench.py > $o/${tag}_microbench.log 2>&1
python tools/sweep.py --sizes 16,18,20 --dists uniform,const,wminus --steps 3 2>&1 | grep "2^" > $o/${tag}_sweep_small.log
python tools/sweep.py --sizes 22,24,26 --dists uniform,const,wminus --steps 3 2>&1 | grep "2^" > $o/${tag}_sweep_large.log
python tools/sweep.py --host --sizes 20,22,24,26 --dists uniform --steps 3 2>&1 | grep "2^" > $o/${tag}_sweep_e2e.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 10 -c 2 -o $o/${tag}_prof_accumulate -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu2.log 2>&1
ls -la $o | tail -12
