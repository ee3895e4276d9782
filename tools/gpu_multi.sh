#!/bin/bash
# multi-GPU validation: the gpu tier on N real GPUs (multi-device tests use them), smoke, and the bench under torchrun
tag=${1:-multi}; n=${2:-2}
o=gpurun_out
mkdir -p $o
nvidia-smi -L | head -8
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -14 $o/${tag}_gpu.log
timeout 300 python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $o/${tag}_smoke.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err; echo "bench rc=$?"; tail -3 $o/${tag}_bench_n$n.err
python -c "
import json
for l in open('$o/${tag}_bench_n$n.json'):
    l = l.strip()
    if not l.startswith('{'): continue
    d = json.loads(l)
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity_vs_oracle'), d['srs_register_ms'])
    print(json.dumps(d.get('strong'))[:1400])
    r = d.get('cojolt_replay')
    if r:
        print(r['srs_generate_register_s'], [(p['party'], p['gpu_seconds'], p['Mpoints_per_s_commit'], p['projected_party_prove_s']) for p in r['parties']])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 2 --warmup 1 2>/dev/null | cut -c1-300
