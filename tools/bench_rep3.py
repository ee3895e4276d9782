#!/usr/bin/env python3
"""Measure the Rep3-polynomial rows on one B200: one JSON line per experiment.

  python tools/bench_rep3.py [--log2n 20] [--k 32] [--nv 18] [--reps 5]

  ingest   cozk_poly_from_wire of one shared polynomial from pinned host memory: H2D GB/s and the canonical->Montgomery
           kernel against the measured HBM peak (algorithmic bytes: 64 B read + 64 B written per coefficient)
  lincomb  cozk_rep3_linear_combination of k resident shared polynomials: bytes/s against the HBM peak and limb
           products/s (2 halves x 64 per term, lazy reduction) against the self-measured integer-multiply peak -
           whichever is closer to 1 is the bound
  chi      cozk_rep3_evaluate_at_chi of the same k polynomials
  open     PST13 opening of a resident polynomial with and without the pair-sum SRS (N1)
"""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6500.0, "fallback 6500 GB/s (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--nv", type=int, default=18)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chi-waves", type=int, default=1)
    ap.add_argument("--bulk", type=int, default=0, help="1: bulk-copy (TMA) staged ingest / chi kernels, 0: plain per-lane loads")
    ap.add_argument("--one-batch-max-nv", type=int, default=-1, help="open_one_batch_max_nv option (-1 default)")
    ap.add_argument("--small-window", type=int, default=0, help="open_small_window option (0 = cost model)")
    ap.add_argument("--small-ragged", type=int, default=-1, help="open_small_ragged option (-1 default)")
    ap.add_argument("--small", default="12", help="comma list of open_small_log2 values to try for the keyed opening")
    args = ap.parse_args()
    cozk = importlib.import_module("co-zkvms_b200")
    rep3, pst = cozk.rep3, cozk.pst13
    ctx = cozk.Context()
    ctx.set_option("bulk_copy", args.bulk)
    ctx.set_option("chi_waves", args.chi_waves)
    if args.one_batch_max_nv >= 0:
        ctx.set_option("open_one_batch_max_nv", args.one_batch_max_nv)
    if args.small_window:
        ctx.set_option("open_small_window", args.small_window)
    if args.small_ragged >= 0:
        ctx.set_option("open_small_ragged", args.small_ragged)
    L = cozk.lib()
    T = cozk.testlib()
    n = 1 << args.log2n
    peak, peak_src = hbm_peak()
    ms, ops = ctx.microbench("fq_mul", 148 * 8, 256, 512)
    imad_peak = ops / (ms * 1e-3) * 136

    # ---- N2 ingest: header | n x 64 random canonical bytes (top byte < 0x30 keeps every value below r) | trailer
    rng = np.random.default_rng(1)
    header = np.frombuffer(int(args.log2n).to_bytes(8, "little") + int(n).to_bytes(8, "little"), np.uint8)
    trailer = np.frombuffer(b"\x00" * 8 + b"\x00" + int(n).to_bytes(8, "little") + b"\x00" * 8 + int(n).to_bytes(8, "little"), np.uint8)
    body = rng.integers(0, 256, size=(2 * n, 32), dtype=np.uint8)
    body[:, 31] &= 0x2F
    pinned = cozk.PinnedBuffer(header.size + body.size + trailer.size)
    pinned.array[:header.size] = header
    pinned.array[header.size:header.size + body.size] = body.reshape(-1)
    pinned.array[header.size + body.size:] = trailer
    h2d, conv, wall = [], [], []
    for _ in range(args.reps + 1):
        t0 = time.perf_counter()
        poly, used = rep3.Rep3DensePolynomial.from_wire(ctx, pinned.array)
        wall.append(time.perf_counter() - t0)
        st = rep3.last_stats(ctx)
        h2d.append(st["h2d_ms"])
        conv.append(st["ingest_ms"])
        poly.release()
    assert used == pinned.nbytes
    h2d_ms, conv_ms = float(np.median(h2d[1:])), float(np.median(conv[1:]))
    print(json.dumps({"experiment": "ingest", "log2n": args.log2n, "wire_bytes": pinned.nbytes, "h2d_ms": h2d_ms,
                      "h2d_gbs": 64 * n / h2d_ms / 1e6, "ingest_kernel_ms": conv_ms,
                      "ingest_kernel_gbs": 128 * n / conv_ms / 1e6, "hbm_peak_gbs": peak, "hbm_peak_source": peak_src,
                      "ingest_frac_of_hbm": 128 * n / conv_ms / 1e6 / peak, "wall_ms": 1e3 * float(np.median(wall[1:]))}))

    # ---- N4: k resident shared polynomials, generated on the device
    polys = []
    tmp = ctx.alloc(n * 64)
    for j in range(args.k):
        for half in (0, 1):
            cozk._check(T.cozk_testgen_scalars(ctx.handle, 0, cozk.DIST["uniform"], 100 + 2 * j + half, 0, n, n, cozk.MONT,
                                               ctypes.c_void_p(tmp.ptr + 32 * half), 64))
        polys.append(rep3.Rep3DensePolynomial.from_device(ctx, tmp, n))
    tmp.free()
    coeffs = np.zeros((args.k, 32), np.uint8)
    coeffs[:] = ctx.testgen_scalars("uniform", 7, args.k).download().reshape(args.k, 32)
    t_ms, nbytes = [], 0
    for _ in range(args.reps + 1):
        ctx.flush_l2()
        joint = rep3.linear_combination(polys, coeffs, 0)
        st = rep3.last_stats(ctx)
        t_ms.append(st["lincomb_ms"])
        nbytes = st["lincomb_bytes"]
        joint.release()
    lin_ms = float(np.median(t_ms[1:]))
    products = 2.0 * 64 * args.k * n
    print(json.dumps({"experiment": "lincomb", "log2n": args.log2n, "k": args.k, "ms": lin_ms, "bytes": nbytes,
                      "gbs": nbytes / lin_ms / 1e6, "frac_of_hbm": nbytes / lin_ms / 1e6 / peak,
                      "limb_products_per_s": products / (lin_ms * 1e-3), "imad_peak": imad_peak,
                      "frac_of_imad": products / (lin_ms * 1e-3) / imad_peak,
                      "terms_per_s": 2.0 * args.k * n / (lin_ms * 1e-3)}))
    chis = ctx.testgen_scalars("uniform", 9, n).download().reshape(n, 32)
    t_ms = []
    for _ in range(args.reps + 1):
        ctx.flush_l2()
        rep3.batch_evaluate_at_chi(polys, chis)
        t_ms.append(rep3.last_stats(ctx)["chi_ms"])
    chi_ms = float(np.median(t_ms[1:]))
    chi_bytes = args.k * n * 64 + n * 32
    print(json.dumps({"experiment": "chi", "log2n": args.log2n, "k": args.k, "ms": chi_ms, "gbs": chi_bytes / chi_ms / 1e6,
                      "frac_of_hbm": chi_bytes / chi_ms / 1e6 / peak,
                      "frac_of_imad": 64.0 * args.k * n / (chi_ms * 1e-3) / imad_peak}))
    # the eq table built on the device instead of copied in: wall clock of the two ways to get evaluate_at_chi's input
    pt = ctx.testgen_scalars("uniform", 13, args.log2n).download().reshape(args.log2n, 32)
    w_host, w_dev, w_eq = [], [], []
    for _ in range(args.reps + 1):
        t0 = time.perf_counter()
        rep3.batch_evaluate_at_chi(polys, chis)
        w_host.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        chi = rep3.eq_evals(ctx, pt, msb_first=True)
        w_eq.append(time.perf_counter() - t0)
        rep3.batch_evaluate_at_chi(polys, chi)
        w_dev.append(time.perf_counter() - t0)
        chi.release()
    print(json.dumps({"experiment": "eq_evals", "log2n": args.log2n, "k": args.k,
                      "eq_evals_wall_ms": 1e3 * float(np.median(w_eq[1:])),
                      "eq_table_gbs": 32 * n / float(np.median(w_eq[1:])) / 1e9,
                      "chi_with_host_table_wall_ms": 1e3 * float(np.median(w_host[1:])),
                      "chi_with_device_eq_table_wall_ms": 1e3 * float(np.median(w_dev[1:]))}))
    for p in polys:
        p.release()

    # ---- N1: opening with / without pair sums
    nv = args.nv
    levels, start = [], 0
    handles = []
    for i in range(nv):
        m = 1 << (nv - i)
        d = ctx.testgen_bases(1, m, start=start)
        handles.append(ctx.srs_register_device(d, m))
        d.free()
        start += m

    class S:
        pass
    setup = S()
    setup.ctx, setup.level_srs = ctx, handles
    m = 1 << nv
    tmp = ctx.testgen_scalars("uniform", 11, m, stride=64)
    poly = rep3.Rep3DensePolynomial.from_device(ctx, tmp, m)
    tmp.free()
    point = ctx.testgen_scalars("uniform", 12, nv).download().reshape(nv, 32)
    def timed_open(keyed):
        ts, proofs = [], None
        for _ in range(args.reps + 1):
            t0 = time.perf_counter()
            proofs, _ = rep3.open_poly(setup, poly, point, keyed=keyed)
            ts.append(time.perf_counter() - t0)
        return 1e3 * float(np.median(ts[1:])), proofs
    ref_ms, ref_proofs = timed_open(False)
    for small in [int(x) for x in args.small.split(",")]:
        ctx.set_option("open_small_log2", small)
        t0 = time.perf_counter()
        rep3.create_open_key(setup)
        key_s = time.perf_counter() - t0
        key_ms, key_proofs = timed_open(True)
        print(json.dumps({"experiment": "open", "nv": nv, "open_small_log2": small, "reference_schedule_ms": ref_ms,
                          "keyed_ms": key_ms, "speedup": ref_ms / key_ms, "open_key_setup_s": key_s,
                          "identical_proofs": bool((ref_proofs == key_proofs).all())}))
        if small == [int(x) for x in args.small.split(",")][-1]:
            # co-spartan's distributed_batch_open_poly_worker on resident share_0 polynomials: aggregate + open + evaluate
            spartan = cozk.spartan
            sp = []
            for j in range(4):
                t = ctx.testgen_scalars("uniform", 30 + j, m)
                sp.append(rep3.Rep3DensePolynomial.from_device(ctx, t, m, kind=rep3.PUBLIC))
                t.free()
            eta = ctx.testgen_scalars("uniform", 14, 1).download().reshape(32)
            ts = []
            for _ in range(args.reps + 1):
                t0 = time.perf_counter()
                spartan.distributed_batch_open_poly_worker(sp, setup, point, eta, 3)
                ts.append(time.perf_counter() - t0)
            print(json.dumps({"experiment": "spartan_batch_open_worker", "nv": nv, "polys": 4, "num_comms": 3,
                              "wall_ms": 1e3 * float(np.median(ts[1:])), "keyed_open_alone_ms": key_ms}))
            for q in sp:
                q.release()
        rep3.release_open_key(setup)
    ctx.close()


if __name__ == "__main__":
    main()
