#!/bin/bash
# Scaling runs of bench.py on one node: weak scaling at 2^20 points per GPU for N = 1, 2, 4, 8 and the sharded sweep of
# BASELINE.json configs[3] (2^24 .. 2^26 points in total across N GPUs).  Usage: tools/run_scaling.sh <max_gpus> <outfile>
MAXG=${1:-8}
OUT=${2:-gpurun_out/scale.jsonl}
: > "$OUT"
run() {  # n_gpus log2n steps
  if [ "$1" -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps "$3" --warmup 3 --log2n "$2" --no-cpu-baseline 2>>"$OUT.err" | grep '^{' >> "$OUT"
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port $((29500 + $1 + $2)) \
      bench.py --gpus "$1" --steps "$3" --warmup 3 --log2n "$2" 2>>"$OUT.err" | grep '^{' >> "$OUT"
  fi
}
for n in 1 2 4 8; do
  [ "$n" -le "$MAXG" ] && run "$n" 20 10
done
for n in 2 4 8; do
  [ "$n" -le "$MAXG" ] || continue
  for total in 24 25 26; do
    case $n in 2) per=$((total - 1));; 4) per=$((total - 2));; 8) per=$((total - 3));; esac
    run "$n" "$per" 3
  done
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print("N=%d 2^%d/GPU  %8.1f Mpts/s  %.2f ms/step  e2e %.1f Mpts/s  c=%s  frac %.2f" % (
        d["n_gpus"], d["config"]["log2_points_per_gpu"], d["value"], d["ms_per_step"], d["e2e"]["value"],
        d["roofline"]["window_bits"], d["roofline"]["pipeline_frac"]))
PY
