#!/usr/bin/env python3
"""SRS registration time (bases on the device -> table of 2^(c w) P rows + row totals), with the two table-build forms:
   python tools/bench_register.py [--sizes 16,20,22,24]"""
import argparse
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,20,22,24")
args = ap.parse_args()
ctx = cozk.Context()
for lg in [int(x) for x in args.sizes.split(",")]:
    n = 1 << lg
    db = ctx.testgen_bases(1, n)
    for rowwise in (1, 0):
        ctx.set_option("table_rowwise", rowwise)
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            srs = ctx.srs_register_device(db, n)
            best = min(best, time.perf_counter() - t0)
            ctx.srs_release(srs)
        print("2^%d points: registration %.3f s (%s)" % (lg, best, "row by row: one inversion per row and point" if rowwise else
                                                          "chain + one inversion per point"), flush=True)
    db.free()
