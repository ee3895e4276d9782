#!/usr/bin/env python3
"""Batched-affine pre-reduction (csrc/affine_kernels.cuh) on its own and inside the MSM.
   python tools/bench_affine.py [--log2n 20]"""
import argparse
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--c", type=int, default=17)
args = ap.parse_args()
ctx = cozk.Context()
n, c = 1 << args.log2n, args.c
W = (254 + 1 + c - 1) // c
# table-mode pair list of a uniform vector: keys = digits (2^(c-1) buckets), vals index a W x n "table" (random points here)
dsc = ctx.testgen_scalars("uniform", 5, n)
keys, vals = ctx.decompose_sort(dsc, n, c, table_stride=n, key_bits=c - 1, fused=True)
dpts = ctx.testgen_bases(7, W * n)
m = keys.size
print("pairs %d, buckets %d, %.1f per bucket" % (m, 1 << (c - 1), m / (1 << (c - 1))))
for rounds in (1, 2, 3, 4):
    best = 1e9
    for _ in range(3):
        _, _, ms = ctx.affine_rounds(keys, vals, dpts, 1 << (c - 1), rounds, download=False)
        best = min(best, ms)
    adds = sum(m >> (r + 1) for r in range(rounds))
    print("rounds %d: %.3f ms, %.1f M additions -> %.2f G additions/s (%.2f T limb products/s at 788 each)" % (
        rounds, best, adds / 1e6, adds / best / 1e6, adds * 788 / best / 1e9))
