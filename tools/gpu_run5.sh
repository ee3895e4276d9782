#!/bin/bash
tag=${1:-r5}
o=gpurun_out
mkdir -p $o
for L in 32 40 48 64; do echo "== acc_chunk $L"; timeout 300 python tools/sweep.py --exact --sizes 20,22,24 --dists uniform --steps 5 --acc-chunk $L 2>&1 | grep "2^"; done | tee $o/${tag}_accl.log
for nv in 16 20 22; do timeout 300 python tools/bench_rep3.py --log2n 18 --k 4 --nv $nv --small 15 2>/dev/null | grep -E '"experiment": "(open|spartan_batch_open_worker)"' | cut -c1-400; done | tee $o/${tag}_open.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-log2n 0 --replay-log2t 0"
$B > $o/${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_accumulate<\(bool\)1>' -s 3 -c 1 -o $o/${tag}_prof_accumulate -f $B > $o/${tag}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $o/${tag}_ncu.log
