#!/usr/bin/env python3
"""Replay of one co-jolt party's commitment path with the share kept in HBM from arrival to opening proof (one process,
--gpus N devices: the polynomials are dealt round-robin to the devices as they arrive, every device commits its own, the
joint polynomial is formed per device and summed on device 0 over NVLink peer mappings, device 0 opens).

What `JoltRep3Prover::init` + `prove` ask of the commitment scheme, through the
resident-polynomial entry points (include/cozk_rep3.h):

  receive   138 shared trace polynomials of length T arrive as ark-serialize images (64 B per coefficient, canonical)
            -> cozk_poly_from_wire: H2D once, canonical -> Montgomery on the device          (witness.rs:130-155)
  commit    PST13::batch_commit_rep3 over the 138 handles (+ 54 final_cts polynomials of 2^16)  (pst13.rs:165-229)
  combine   Rep3MultilinearPolynomial::linear_combination of all of them with powers of gamma   (opening_proof.rs:268-278)
  open      PST13::prove_rep3 of the joint polynomial with the opening key                      (pst13.rs:125-137)

One pinned wire image per share shape is reused for every polynomial of that shape: PCIe traffic and arithmetic are those
of distinct polynomials, only host RAM is saved.  Prints one JSON line per (T, party) next to the reference's
trace-derived CPU seconds.  The Rust prover cannot run here, so this is a replay, not a proof.
"""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
cozk = importlib.import_module("co-zkvms_b200")
rep3, pst = cozk.rep3, cozk.pst13

REF = {16: (32, 11.18, 5.53, 0.09, 1.37), 20: (8, 204.6, 152.0, 3.40, 71.3), 22: (32, 345.9, 145.8, 5.57, 325.5)}
N_SHARED, N_FINAL = 138, 54


def wire_image(ctx, dist, seed, n):
    """Pinned ark-serialize image of a Rep3DensePolynomial of n coefficients: share a ~ dist, share b uniform."""
    L = cozk.lib()
    T = cozk.testlib()
    d = ctx.alloc(n * 64)
    for half, (dd, sd) in enumerate(((dist, seed), ("uniform", seed + 1000))):
        cozk._check(T.cozk_testgen_scalars(ctx.handle, 0, cozk.DIST[dd], sd, 0, n, n, cozk.CANON,
                                           ctypes.c_void_p(d.ptr + 32 * half), 64))
    nv = n.bit_length() - 1
    header = nv.to_bytes(8, "little") + n.to_bytes(8, "little")
    trailer = (0).to_bytes(8, "little") + b"\x00" + n.to_bytes(8, "little") + (0).to_bytes(8, "little") + n.to_bytes(8, "little")
    buf = cozk.PinnedBuffer(len(header) + n * 64 + len(trailer))
    buf.array[:16] = np.frombuffer(header, np.uint8)
    buf.array[16:16 + n * 64] = d.download()
    buf.array[16 + n * 64:] = np.frombuffer(trailer, np.uint8)
    d.free()
    return buf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2t", default="16")
    ap.add_argument("--parties", default="0,2")
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    G = args.gpus
    ctx = cozk.Context(devices=list(range(G)))
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(G)

    def receive(image, count):
        """`count` polynomials from one wire image, dealt round-robin; one host thread per device (ctypes drops the GIL)."""
        def on_device(d):
            return [rep3.Rep3DensePolynomial.from_wire(ctx, image.array, device=d)[0] for _ in range(d, count, G)]
        return [p for part in pool.map(on_device, range(G)) for p in part]

    dist_of_party = {0: "const", 1: "const", 2: "wminus"}
    for lt in [int(x) for x in args.log2t.split(",")]:
        T = 1 << lt
        n16 = min(T, 1 << 16)
        levels, start = [], 0
        t0 = time.perf_counter()
        for i in range(lt):
            n = 1 << (lt - i)
            d = ctx.testgen_bases(1, n, start=start)
            levels.append(ctx.srs_register_device(d, n))
            d.free()
            start += n
        setup = pst.PST13Setup.__new__(pst.PST13Setup)
        setup.ctx, setup.level_srs, setup.num_vars = ctx, levels, lt
        rep3.create_open_key(setup)
        setup_s = time.perf_counter() - t0
        gamma = ctx.testgen_scalars("uniform", 5, N_SHARED + N_FINAL)
        coeffs = gamma.download().reshape(-1, 32)
        gamma.free()
        dpt = ctx.testgen_scalars("uniform", 7, lt)
        point = dpt.download().reshape(lt, 32)
        dpt.free()
        for party in [int(x) for x in args.parties.split(",")]:
            img = wire_image(ctx, dist_of_party[party], 100 + party, T)
            img16 = wire_image(ctx, dist_of_party[party], 200 + party, n16)
            times = {}
            for rep in range(2):  # pass 0 warms the engine's scratch buffers up to their final size
                t0 = time.perf_counter()
                polys = receive(img, N_SHARED)
                finals = receive(img16, N_FINAL)
                times["receive_s"] = time.perf_counter() - t0
                t0 = time.perf_counter()
                comms = rep3.batch_commit_rep3(setup, polys, commit_to_public=False)
                times["commit_trace_polys_s"] = time.perf_counter() - t0
                t0 = time.perf_counter()
                rep3.batch_commit_rep3(setup, finals, commit_to_public=False)  # against the 2^16 prefix of the same SRS
                times["commit_final_cts_s"] = time.perf_counter() - t0
                t0 = time.perf_counter()
                joint = rep3.linear_combination(polys + finals, coeffs, party)
                times["linear_combination_s"] = time.perf_counter() - t0
                lin_stats = rep3.last_stats(ctx)
                t0 = time.perf_counter()
                proofs, _ = rep3.prove_rep3(setup, joint, point)
                times["prove_rep3_s"] = time.perf_counter() - t0
                for p in polys + finals + [joint]:
                    p.release()
            total = sum(times.values())
            wire_bytes = N_SHARED * img.nbytes + N_FINAL * img16.nbytes
            line = {"config": "co-jolt party commitment path, share resident in HBM", "log2_T": lt, "party": party,
                    "share_dist": dist_of_party[party], "gpus": G,
                    "partial_sum_kernel_ms": round(lin_stats["partial_sum_ms"], 3), "peer_gbytes": round(lin_stats["peer_bytes"] / 1e9, 3), "gpu_seconds": {k: round(v, 4) for k, v in times.items()},
                    "total_s": round(total, 4), "wire_gbytes": round(wire_bytes / 1e9, 3),
                    "receive_gbytes_per_s": round(wire_bytes / times["receive_s"] / 1e9, 1),
                    "commit_Mpoints_per_s": round(N_SHARED * T / times["commit_trace_polys_s"] / 1e6, 1),
                    "setup_s_srs_tables_open_key": round(setup_s, 3),
                    "x_first_commitment": bytes(comms[0].g_product[:6]).hex(), "x_first_proof": bytes(proofs[0][:6]).hex()}
            if lt in REF:
                vcpu, prove, msm_ref, open_ref, init_ref = REF[lt]
                line["reference_cpu"] = {"vcpu": vcpu, "party_prove_s": prove, "batch_msm_s": msm_ref, "prove_rep3_s": open_ref,
                                         "init_receive_s": init_ref, "source": "co-jolt/traces (BASELINE.md 1.1), party 0"}
            print(json.dumps(line), flush=True)
            img.free()
            img16.free()
        rep3.release_open_key(setup)
        for h in levels:
            ctx.srs_release(h)
    ctx.close()


if __name__ == "__main__":
    main()
