#!/bin/bash
tag=${1:-r8}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "streamed or dominant or sliced" > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -5 $o/${tag}_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --strong-log2n 0 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['e2e']['value']); print([(p['party'], p['gpu_seconds']) for p in d['cojolt_replay']['parties']])"
