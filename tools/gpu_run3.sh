#!/bin/bash
tag=${1:-r3}
o=gpurun_out
mkdir -p $o
for b in 1 0; do for lg in 20 22; do
  echo "== bulk=$b log2n=$lg"; timeout 300 python tools/bench_rep3.py --bulk $b --log2n $lg --k 32 --nv 16 2>/dev/null | grep -E '"experiment": "(ingest|chi|lincomb)"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ('experiment', 'ingest_kernel_ms', 'ingest_frac_of_hbm', 'ms', 'frac_of_hbm', 'frac_of_imad', 'gbs')})"
done; done | tee $o/${tag}_bulk.log
timeout 900 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"; tail -3 $o/${tag}_bench.err
python -c "
import json; d=json.load(open('$o/${tag}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'], d.get('parity_vs_oracle'))
print(json.dumps(d.get('strong'))[:1500])
print(json.dumps(d.get('cojolt_replay'))[:3000])
print(d.get('cpu_baseline'))"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_ref.json 2>> $o/${tag}_bench.err; echo "ref rc=$?"; cut -c1-400 $o/${tag}_bench_ref.json
