#!/bin/bash
# final check of a build: whole gpu tier, smoke, the driver's bench line and the reference arm
tag=${1:-final}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -4 $o/${tag}_gpu.log
timeout 300 python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $o/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench_1gpu.json 2> $o/${tag}_bench.err; echo "bench rc=$?"; tail -3 $o/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_bench.err; echo "ref rc=$?"
python -c "
import json; d=json.load(open('$o/${tag}_bench_1gpu.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'], d.get('parity_vs_oracle'))
s=d['strong']; print({k:s[k] for k in ('e2e_ms','device_ms_max','parity_vs_oracle','srs_register_ms')})
print([(p['party'],p['gpu_seconds']) for p in d['cojolt_replay']['parties']])
print(d['cpu_baseline']['value'], json.load(open('$o/${tag}_bench_reference_arm.json'))['value'])"
