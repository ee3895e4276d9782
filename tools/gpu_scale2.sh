#!/bin/bash
# strong record on N GPUs with different streaming thresholds of the sliced path
tag=${1:-scale2}; n=${2:-8}
o=gpurun_out
mkdir -p $o
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:2}"; }
for opt in "stream_min_points_sliced=4194304" "stream_min_points_sliced=2097152,stream_chunks=2" "stream_min_points_sliced=2097152,stream_chunks=4" "stream_min_points_sliced=1048576,stream_chunks=2"; do
  timeout 600 bash -c "$(declare -f run); n=$n; run 29531 --steps 3 --warmup 3 --strong-steps 8 --replay-log2t 0 --no-cpu-baseline --strong-opts $opt" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        s = json.loads(l)['strong']
        print(s.get('options'), {k: round(s[k], 3) for k in ('e2e_ms', 'device_ms_max', 'compute_ms_max', 'e2e_mpoints_per_s')})"
done | tee $o/${tag}_stream.log
