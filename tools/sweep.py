#!/usr/bin/env python3
"""Size x distribution sweep of the single-GPU MSM (device-resident scalars), with stage breakdown.
   python tools/sweep.py [--sizes 16,18,20,22,24] [--dists uniform,const,wminus] [--steps 3] [--batch 1]"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,18,20,22,24")
ap.add_argument("--dists", default="uniform,const,wminus")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--window", type=int, default=0)
ap.add_argument("--stride", type=int, default=32)
ap.add_argument("--bits", type=int, default=0)
ap.add_argument("--stream-min", type=int, default=-1, help="stream_min_points option (-1 default, 0 never)")
ap.add_argument("--stream-chunks", type=int, default=0)
ap.add_argument("--table-window", type=int, default=-1, help="-1 auto, 0 no table, else forced table window")
ap.add_argument("--acc-chunk", type=int, default=0, help="pairs per level-1 accumulate thread (0 = chosen per call)")
ap.add_argument("--acc-chunk-up", type=int, default=0, help="partial slots per thread at the serial accumulate levels >= 2")
ap.add_argument("--sort-bits", type=int, default=0, help="digit bits per pass of the pair sort (7 .. 11; default 8)")
ap.add_argument("--group-l", type=int, default=0, help="buckets per thread in the group step of the bucket reduce")
ap.add_argument("--exact", action="store_true", help="register an SRS of exactly each size (the bench's shape) instead of prefixes of the largest")
ap.add_argument("--dominant", type=int, default=-1, help="-1 default, 0 off, 1 on (dominant-digit mode of whole-SRS calls)")
ap.add_argument("--stream-first-pct", type=int, default=0)
ap.add_argument("--chunk-min", type=int, default=-1, help="chunk_min_points option: device-resident vectors this long run in chunks (0 never)")
ap.add_argument("--reduce-2d", type=int, default=-1, help="bucket reduce: 1 row / column form, 0 group form (-1 default)")
ap.add_argument("--affine-rounds", type=int, default=-1, help="batched-affine pre-reduction rounds (-1 default, 0 off)")
ap.add_argument("--host", action="store_true", help="scalars in pinned host memory (end to end)")
args = ap.parse_args()

ctx = cozk.Context()
if args.window:
    ctx.set_option("window", args.window)
if args.dominant >= 0:
    ctx.set_option("dominant", args.dominant)
if args.acc_chunk:
    ctx.set_option("acc_chunk", args.acc_chunk)
if args.acc_chunk_up:
    ctx.set_option("acc_chunk_up", args.acc_chunk_up)
if args.sort_bits:
    ctx.set_option("sort_digit_bits", args.sort_bits)
if args.group_l:
    ctx.set_option("group_l", args.group_l)
if args.chunk_min >= 0:
    ctx.set_option("chunk_min_points", args.chunk_min)
if args.stream_first_pct:
    ctx.set_option("stream_first_pct", args.stream_first_pct)
if args.reduce_2d >= 0:
    ctx.set_option("reduce_2d", args.reduce_2d)
if args.affine_rounds >= 0:
    ctx.set_option("affine_rounds", args.affine_rounds)
if args.stream_min >= 0:
    ctx.set_option("stream_min_points", args.stream_min)
if args.stream_chunks:
    ctx.set_option("stream_chunks", args.stream_chunks)
if args.table_window == 0:
    ctx.set_option("table_max_mib", 0)
elif args.table_window > 0:
    ctx.set_option("table_window", args.table_window)
sizes = [int(x) for x in args.sizes.split(",")]
nmax = 1 << max(sizes)
srs = None
if not args.exact:
    dbases = ctx.testgen_bases(1, nmax)
    srs = ctx.srs_register_device(dbases, nmax)
    dbases.free()
rows = []
for dist in args.dists.split(","):
    ds = [ctx.testgen_scalars(dist, 2 + j, nmax, stride=args.stride) for j in range(args.batch)]
    if args.host:
        pins = []
        for d in ds:
            pb = cozk.PinnedBuffer(nmax * args.stride)
            pb.array[:] = d.download()
            pins.append(pb)
    for lg in sizes:
        n = 1 << lg
        if args.exact:
            if srs is not None:
                ctx.srs_release(srs)
            db = ctx.testgen_bases(1, n)
            srs = ctx.srs_register_device(db, n)
            db.free()
        out = np.zeros((args.batch, 72), np.uint8)
        ptrs = [pb.ptr for pb in pins] if args.host else [d.ptr for d in ds]
        dev = None if args.host else 0
        form = 1 if args.bits else 0
        for _ in range(2):
            ctx.msm_batch_ptrs(srs, ptrs, n, stride=args.stride, device=dev, out=out, max_num_bits=args.bits, form=form)
        acc = {}
        for _ in range(args.steps):
            ctx.flush_l2()
            ctx.msm_batch_ptrs(srs, ptrs, n, stride=args.stride, device=dev, out=out, max_num_bits=args.bits, form=form)
            st = ctx.last_stats()
            for k, v in st.items():
                acc[k] = acc.get(k, 0.0) + v
        st = {k: v / args.steps for k, v in acc.items()}
        row = {"dist": dist, "log2n": lg, "batch": args.batch, "ms": st["total_ms"],
               "Mpts_s": args.batch * n / st["total_ms"] / 1e3, "c": int(st["window"]), "W": int(st["windows"]),
               "stages": {k: round(st[k], 3) for k in ("decompose_ms", "sort_ms", "accumulate_ms", "reduce_ms", "finish_ms", "h2d_ms")},
               "x": bytes(out[0][:6]).hex()}
        rows.append(row)
        print("%-8s 2^%-2d k=%-3d c=%-2d W=%-2d %9.3f ms %8.1f Mpts/s  dec %.2f sort %.2f acc %.2f red %.2f fin %.2f other %.2f" % (
            dist, lg, args.batch, row["c"], row["W"], row["ms"], row["Mpts_s"], st["decompose_ms"], st["sort_ms"],
            st["accumulate_ms"], st["reduce_ms"], st["finish_ms"], st["h2d_ms"]), flush=True)
    for d in ds:
        d.free()
print(json.dumps(rows))
