#!/bin/bash
# bench under torchrun on N GPUs: the driver's line (weak scaling + strong 2^24 record + co-jolt replay), then the 2^26 strong record
tag=${1:-scale}; n=${2:-8}
o=gpurun_out
mkdir -p $o
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:2}"; }
timeout 1500 bash -c "$(declare -f run); n=$n; run 29521 --steps 10 --warmup 3" > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err; echo "bench rc=$?"; tail -2 $o/${tag}_bench_n$n.err
timeout 1500 bash -c "$(declare -f run); n=$n; run 29522 --steps 5 --warmup 3 --strong-log2n 26 --strong-steps 3 --replay-log2t 0" > $o/${tag}_bench_n${n}_s26.json 2> $o/${tag}_bench_n${n}_s26.err; echo "bench26 rc=$?"; tail -2 $o/${tag}_bench_n${n}_s26.err
python - <<PY
import json
for f in ("$o/${tag}_bench_n$n.json", "$o/${tag}_bench_n${n}_s26.json"):
    for l in open(f):
        l = l.strip()
        if not l.startswith('{'): continue
        d = json.loads(l)
        print(f, d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 3), round(d['e2e']['value'], 1), d.get('parity_vs_oracle'), round(d['srs_register_ms'], 1))
        s = d.get('strong')
        if s: print({k: s[k] for k in ('log2_points_total', 'n_devices', 'srs_register_ms', 'e2e_ms', 'e2e_mpoints_per_s', 'device_ms_max', 'device_mpoints_per_s', 'compute_ms_max', 'window_bits', 'windows', 'parity_vs_oracle') if k in s})
        r = d.get('cojolt_replay')
        if r: print(r['srs_generate_register_s'], [(p['party'], p['gpu_seconds'], p['Mpoints_per_s_commit'], p['projected_party_prove_s']) for p in r['parties']])
PY
