#!/usr/bin/env python3
"""Replay of the co-jolt party MSM schedule (BASELINE.json configs[0] and configs[4]) through the reference-facing API.

The Rust prover cannot run here (no toolchain), so this replays what one party's `JoltRep3Prover::prove` asks of the
commitment scheme with synthetic shares of the real shapes:

  commit phase   PST13::batch_commit_rep3(read_write_values)   138 shared (AoS {a,b}, stride 64) + 60 public polys, len T
                 PST13::commit_rep3 x 3                        t_final (public), v_final, t_final (len T)
                 PST13::batch_commit_rep3(final_cts)           54 shared polys, len 2^16
  opening        PST13::prove_rep3 -> open()                   nv MSMs of sizes 2^nv .. 2 with duplicated scalars

Share distributions per party (co-jolt/src/poly/dense_mlpoly.rs:553-588): party 0 and 1 hold a CONSTANT vector, party 2
holds w - c0 - c1.  One host buffer per distribution is reused for all polynomials of that shape (the arithmetic and
the PCIe traffic are those of distinct polynomials; only host RAM is saved).  Host buffers are pinned.

Prints one JSON line per (T, party) with GPU seconds per phase next to the reference's trace-derived CPU seconds
 and the projected party prove time = reference non-MSM time + measured GPU MSM time.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")
pst = cozk.pst13

# reference numbers, party 0, from the bundled traces: T -> (vCPU, prove s, batch_msm s, prove_rep3 s)
REF = {16: (32, 11.18, 5.53, 0.09), 20: (8, 204.6, 152.0, 3.40), 22: (32, 345.9, 145.8, 5.57)}
N_SHARED, N_PUBLIC, N_FINAL = 138, 60, 54
PUBLIC_BITS = [1] * 40 + [16] * 10 + [32] * 5 + [64] * 5   # flags / u16 / u32 / u64 coefficient polynomials (inferred split)


def make_setup(ctx, lt, with_open=True):
    """SRS of a 2^lt-cycle proof: level i holds 2^(lt - i) points (pst13.rs:233-250); generated on the device, registered per
    level.  Returns (setup, seconds)."""
    levels, start = [], 0
    t0 = time.perf_counter()
    for i in range(lt if with_open else 1):
        n = 1 << (lt - i)
        d = ctx.testgen_bases(1, n, start=start)
        levels.append(ctx.srs_register_device(d, n))
        d.free()
        start += n
    setup = pst.PST13Setup.__new__(pst.PST13Setup)
    setup.ctx, setup.level_srs, setup.num_vars = ctx, levels, lt
    return setup, time.perf_counter() - t0


def replay_party(ctx, setup, lt, party, do_open=True, gpus=1, srs_s=0.0):
    """One party's commitment-path calls for a 2^lt-cycle trace; returns the JSON-able record."""
    rep3 = cozk.rep3
    T = 1 << lt
    dist = {0: "const", 1: "const", 2: "wminus"}[party]
    seed = 100 + party
    # one pinned AoS share buffer of length T (a = the share, b = filler) and the public polynomials in their PACKED forms
    # (MultilinearPolynomial::U8Scalars .. U64Scalars, multilinear_polynomial.rs:226-268): 1 - 8 bytes per coefficient over PCIe
    aos = cozk.PinnedBuffer(T * 64)
    d = ctx.testgen_scalars(dist, seed, T, stride=64)
    aos.array[:] = d.download()
    d.free()
    shared = aos.array.reshape(T, 64)
    kind_of = {1: (rep3.U8, np.uint8), 16: (rep3.U16, np.uint16), 32: (rep3.U32, np.uint32), 64: (rep3.U64, np.uint64)}
    pubs = {}
    for bits in sorted(set(PUBLIC_BITS)):
        kd, dt = kind_of[bits]
        pb = cozk.PinnedBuffer(T * np.dtype(dt).itemsize)
        vals = np.random.default_rng(bits).integers(0, 2 if bits == 1 else (1 << min(bits, 63)), size=T, dtype=np.uint64).astype(dt)
        pb.array[:] = vals.view(np.uint8)
        pubs[bits] = (pb, kd, dt)
    n_pub = N_PUBLIC if party == 0 else 0   # parties 1/2 drop the public results; the drop-in skips those MSMs
    pub_polys = [pubs[b][0].array.view(pubs[b][2]) for b in PUBLIC_BITS[:n_pub]]
    pub_kinds = [pubs[b][1] for b in PUBLIC_BITS[:n_pub]]
    times = {}
    # untimed warm-up with the real shapes: the engine's scratch buffers grow to their final size here
    pst.batch_commit_rep3(setup, [shared] * N_SHARED, [True] * N_SHARED, commit_to_public=False)
    if n_pub:
        pst.batch_commit_packed(setup, pub_polys, pub_kinds, commit_to_public=True)
    t0 = time.perf_counter()
    out = pst.batch_commit_rep3(setup, [shared] * N_SHARED, [True] * N_SHARED, commit_to_public=False)
    times["trace_polys_shared_s"] = time.perf_counter() - t0
    if n_pub:
        t0 = time.perf_counter()
        pst.batch_commit_packed(setup, pub_polys, pub_kinds, commit_to_public=True)
        times["trace_polys_public_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(2):
        pst.batch_commit_rep3(setup, [shared], [True], commit_to_public=False)     # v_final, t_final
    if party == 0:
        pst.batch_commit_packed(setup, [pubs[32][0].array.view(np.uint32)], [rep3.U32])  # bytecode t_final (public)
    times["single_commits_s"] = time.perf_counter() - t0
    n16 = min(T, 1 << 16)
    t0 = time.perf_counter()
    pst.batch_commit_rep3(setup, [shared[:n16]] * N_FINAL, [True] * N_FINAL, commit_to_public=False)
    times["final_cts_s"] = time.perf_counter() - t0
    if do_open:
        point = np.zeros((lt, 32), np.uint8)
        dpt = ctx.testgen_scalars("uniform", 7, lt)
        point[:] = dpt.download().reshape(lt, 32)
        dpt.free()
        t0 = time.perf_counter()
        pst.open(setup, shared, point, stride=64)
        times["open_s"] = time.perf_counter() - t0
    msm_gpu = sum(v for k, v in times.items() if k != "open_s")
    vcpu, prove, msm_ref, open_ref = REF.get(lt, (None, None, None, None))
    line = {"config": "co-jolt party MSM replay", "log2_T": lt, "party": party, "share_dist": dist, "gpus": gpus,
            "gpu_seconds": {k: round(v, 4) for k, v in times.items()},
            "gpu_commit_msm_s": round(msm_gpu, 4), "gpu_open_s": round(times.get("open_s", 0.0), 4),
            "large_scalar_point_mults": N_SHARED * T + N_FINAL * n16 + 2 * T,
            "Mpoints_per_s_commit": round((N_SHARED * T + N_FINAL * n16 + 2 * T) / msm_gpu / 1e6, 1),
            "srs_generate_register_s": round(srs_s, 3),
            "public_polys": "%d in packed form (u8 / u16 / u32 / u64), widened on the device" % n_pub,
            "x_first_commitment": bytes(out[0].g_product[:6]).hex()}
    if prove is not None:
        non_msm = prove - msm_ref - open_ref
        line["reference_cpu"] = {"vcpu": vcpu, "party_prove_s": prove, "batch_msm_s": msm_ref, "prove_rep3_s": open_ref,
                                 "source": "co-jolt/traces (BASELINE.md 1.1), party 0"}
        line["projected_party_prove_s"] = round(non_msm + msm_gpu + times.get("open_s", 0.0), 2)
        line["projection_note"] = "reference non-MSM time from the trace + measured GPU MSM/open time; the Rust prover cannot run here"
    aos.free()
    for pb, _, _ in pubs.values():
        pb.free()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2t", default="16")
    ap.add_argument("--parties", default="0,1,2")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-open", action="store_true")
    args = ap.parse_args()
    ctx = cozk.Context(devices=list(range(args.gpus)))
    for lt in [int(x) for x in args.log2t.split(",")]:
        setup, srs_s = make_setup(ctx, lt, with_open=not args.no_open)
        for party in [int(x) for x in args.parties.split(",")]:
            print(json.dumps(replay_party(ctx, setup, lt, party, do_open=not args.no_open, gpus=args.gpus, srs_s=srs_s)), flush=True)
        for h in setup.level_srs:
            ctx.srs_release(h)
    ctx.close()


if __name__ == "__main__":
    main()
