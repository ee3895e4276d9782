#!/bin/bash
tag=${1:-affp}
o=gpurun_out
mkdir -p $o
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_affine --csv --log-file $o/${tag}_launches.csv python tools/bench_affine.py --log2n 20 > $o/${tag}_ncu1.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('$o/${tag}_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:40]:
    print(r[ki][:40], r[vi])
PY
ncu --set full --clock-control none --import-source on -k regex:k_affine_apply -c 1 -o $o/${tag}_apply -f python tools/bench_affine.py --log2n 20 > $o/${tag}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_affine_prod -c 1 -o $o/${tag}_prod -f python tools/bench_affine.py --log2n 20 > $o/${tag}_ncu3.log 2>&1
ls -la $o | tail -5
