#!/bin/bash
tag=${1:-open4}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_rep3.py tests/test_gpu_pst13.py -x -q -m gpu > $o/${tag}_gpu.log 2>&1; echo "gpu rc=$?"; tail -3 $o/${tag}_gpu.log
for cfg in "12 15" "16 15" "18 15" "20 15" "21 15" "22 15"; do
  set -- $cfg
  timeout 600 python tools/bench_rep3.py --log2n 16 --k 2 --nv $1 --small $2 2>/dev/null | grep '"open"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('nv', d['nv'], 'small_log2', d['open_small_log2'], 'keyed_ms %.3f' % d['keyed_ms'], 'reference_ms %.3f' % d['reference_schedule_ms'], 'setup_s %.2f' % d['open_key_setup_s'], d['identical_proofs'])"
done | tee $o/${tag}_open.log
