#!/usr/bin/env python3
"""Integer-pipe and field-arithmetic microbenchmarks on cuda:0 (self-measured roofline denominators)."""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")

ctx = cozk.Context()
res = {}
for name, blocks, threads, iters in [("imad", 148 * 8, 256, 4096), ("imad_cc", 148 * 8, 256, 2048), ("imad_lo", 148 * 8, 256, 4096),
                                     ("imad_hi", 148 * 8, 256, 4096), ("fq_mul", 148 * 8, 256, 512), ("fq_mul4", 148 * 6, 256, 256),
                                     ("fq_sqr", 148 * 8, 256, 512), ("madd", 148 * 16, 128, 256),
                                     ("fq_mul", 148 * 2, 128, 512), ("fq_mul", 148 * 4, 128, 512), ("fq_mul", 148 * 4, 256, 512),
                                     ("madd", 148 * 4, 128, 256), ("madd", 148 * 8, 128, 256)]:
    ms, ops = ctx.microbench(name, blocks, threads, iters)
    key = "%s[%dx%d]" % (name, blocks // 148, threads)
    res[key] = {"ms": ms, "Gops_per_s": ops / ms / 1e6}
    print(key, "%.3f ms  %.1f Gops/s" % (ms, ops / ms / 1e6))
print(json.dumps(res))
