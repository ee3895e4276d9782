#!/usr/bin/env python3
"""SRS generation, G1 half: time cozk_fixed_base_batch_mul for the 2^(nv+1) - 2 eq-basis scalars of
PST13::setup and print it next to the reference's `PST13::setup` span (co-jolt/traces; includes the G2 half there)."""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")
REF = {16: (0.98, 32), 20: (18.5, 8), 22: (25.7, 32)}
ctx = cozk.Context()
base = np.zeros(72, np.uint8)
one = 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d  # R mod p: Montgomery 1; generator (1, 2)
base[:32] = np.frombuffer(one.to_bytes(32, "little"), np.uint8)
base[32:64] = np.frombuffer(((2 * one) % 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47).to_bytes(32, "little"), np.uint8)
for nv in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,20").split(",")]:
    n = (1 << (nv + 1)) - 2
    d = ctx.testgen_scalars("uniform", 5, n)
    sc = d.download().reshape(n, 32)
    d.free()
    ctx.fixed_base_batch_mul(base, sc[:1024])  # warm-up
    t0 = time.perf_counter()
    pts, _ = ctx.fixed_base_batch_mul(base, sc)
    dt = time.perf_counter() - t0
    ref = REF.get(nv)
    print(json.dumps({"op": "fixed-base batch mul (SRS generation, G1 half)", "nv": nv, "scalars": n, "seconds_end_to_end": round(dt, 4),
                      "Mpoints_per_s": round(n / dt / 1e6, 1), "x0": bytes(pts[0][:6]).hex(),
                      "reference_PST13_setup_s": ref[0] if ref else None, "reference_vcpu": ref[1] if ref else None,
                      "note": "reference span covers G1 and G2 halves plus the eq-basis scalars"}), flush=True)
