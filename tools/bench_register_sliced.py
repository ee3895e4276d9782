#!/usr/bin/env python3
"""Sliced SRS registration over all visible GPUs (one process): python tools/bench_register_sliced.py [--log2n 24]"""
import argparse
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=24)
args = ap.parse_args()
n = 1 << args.log2n
nd = torch.cuda.device_count()
with cozk.Context(devices=list(range(nd))) as ctx:
    db = ctx.testgen_bases(1, n)
    bases = db.download().reshape(n, 64)
    db.free()
    for rowwise in (0, 1, 0):
        ctx.set_option("table_rowwise", rowwise)
        t0 = time.perf_counter()
        srs = ctx.srs_register(bases, sliced=True)
        dt = time.perf_counter() - t0
        ctx.srs_release(srs)
        print("%d GPUs, 2^%d points sliced: registration %.3f s (%s)" % (nd, args.log2n, dt, "row by row" if rowwise else "chain + one inversion per point"), flush=True)
