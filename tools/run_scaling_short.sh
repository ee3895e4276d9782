#!/bin/bash
# The three 8-GPU lines that matter most (a full tools/run_scaling.sh costs ~35 GPU-minutes): weak scaling at 2^20 per
# GPU, the north-star 2^24 points on 8 GPUs, and 2^26 points on 8 GPUs.  Usage: tools/run_scaling_short.sh <gpus> <outfile>
N=${1:-8}
OUT=${2:-gpurun_out/scale_short.jsonl}
: > "$OUT"
for lg in 20 21 23; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + lg)) \
    bench.py --gpus "$N" --steps $([ $lg -le 21 ] && echo 10 || echo 3) --warmup 3 --log2n "$lg" 2>>"$OUT.err" | grep '^{' >> "$OUT"
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print("N=%d 2^%d/GPU  %8.1f Mpts/s  %.2f ms/step  e2e %.1f Mpts/s  c=%s  frac %.2f  clocks %s" % (
        d["n_gpus"], d["config"]["log2_points_per_gpu"], d["value"], d["ms_per_step"], d["e2e"]["value"],
        d["roofline"]["window_bits"], d["roofline"]["pipeline_frac"], d["clocks"]))
PY
