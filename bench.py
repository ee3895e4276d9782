#!/usr/bin/env python3
"""bench.py - BN254 G1 MSM throughput (BASELINE.json metric) on N GPUs of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log2n 20] [--dist uniform] [--impl ours|reference]

A step is one MSM of 2^log2n points per GPU (config "standalone BN254 G1 MSM 2^20 random share scalars on 1xB200",
BASELINE.json configs[1]).  With N > 1 (torchrun, one rank per GPU) the global MSM has N * 2^log2n points, rank r owns
the point range [r, r+1) * 2^log2n (weak scaling), and the N partial sums are added on the host of rank 0 - no data-path
collective.

value   : whole-job Mpoints/s with scalars and bases resident in HBM (cozk_msm_batch_device), timed with the engine's
          CUDA events on its launching stream, max over ranks; L2 flushed between steps.
e2e     : the same MSM through the reference-facing call with HOST scalars in pinned memory (cozk_msm_batch):
          H2D of 32 B/point and D2H of the 72-byte result are inside the timed region (wall clock around the call).
roofline: the dominant kernel (bucket accumulation) against the SELF-MEASURED integer-multiply pipe peak
          (MEASURED_PEAKS.json has no integer figure); HBM traffic is reported to show it is not the bound.
cpu_baseline: the CPU oracle (C restatement of arkworks' Pippenger; the Rust reference cannot be built here) on the
          box's host cores.  --impl reference prints only that, as its own line.
strong  : BASELINE.json configs[3] through the library's own single-process multi-device path (the Rust caller's shape):
          ONE context over all N GPUs on rank 0, the SRS registered in slices (each GPU holds only its point range and its
          table), ONE 2^24-point MSM from pinned host scalars through cozk_msm_batch, partial sums combined inside the call;
          the result is compared with the CPU oracle.  The other ranks idle meanwhile.  --strong-log2n 0 skips it.
parity_vs_oracle: the (combined) point of the timed MSM against the CPU oracle, at every N.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "BN254 G1 MSM Mpoints/s"
CANON_MULTS_PER_POINT = 160          # c = 16, W = 16, 10 field mults per mixed add
LIMB_PRODUCTS_PER_MULT = 136         # 8x8 + 8x8 + 8 32-bit limb products per Montgomery multiplication
# what the kernels actually issue per group addition (SASS-verified counts, tests/test_field_ptx.py): a mixed addition is
# 6 multiplications (136) + 2 squarings (108) + 1 fused product pair (200); a full addition 10 x 136 + 2 x 108 + 200
LIMB_PRODUCTS_PER_MADD = 6 * 136 + 2 * 108 + 200


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region: NVML (sub-millisecond per query, so that even a
    35 ms timed region gets dozens of samples), nvidia-smi as the fallback."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        return [str(sm), str(self.max_mhz), "0"] + ["Active" if f else "Not Active" for f in flags]

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                self.nvml = None
            time.sleep(0.02)  # an nvidia-smi query takes ~30-50 ms itself: about 15 samples per second

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 7 and s[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_reference_run(log2n, dist, steps, warmup, threads):
    """The reference's CPU algorithm (oracle restatement) on the host cores; returns (Mpoints/s, ms/step, sample text)."""
    from oracle import orc
    orc.build()
    n = 1 << log2n
    bases = orc.gen_bases(1, n, threads=threads)
    sc = orc.gen_scalars(dist, 2, n)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        orc.msm(bases, sc, threads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return n * len(times) / total / 1e6, 1e3 * total / len(times), "2^%d points, %s scalars, %d threads" % (log2n, dist, threads)


def strong_record(cozk, args, n_devices, check):
    """BASELINE.json configs[3]: one MSM of 2^strong_log2n points sharded by point range over all GPUs of the box, driven
    by ONE process through cozk_msm_batch with host scalars (what the Rust caller does); sliced SRS."""
    n = 1 << args.strong_log2n
    rec = {"log2_points_total": args.strong_log2n, "n_devices": n_devices, "scalar_dist": args.dist,
           "path": "one cozk_ctx over all devices, cozk_srs_register_sliced, cozk_msm_batch(host scalars, k=1); partial sums "
                   "added on the host inside the timed call"}
    with cozk.Context(devices=list(range(n_devices))) as mctx:
        for kv in [x for x in args.strong_opts.split(",") if x]:
            name, val = kv.split("=")
            mctx.set_option(name, int(val))
            rec.setdefault("options", {})[name] = int(val)
        gb = mctx.testgen_bases(1, n)
        hb = gb.download().reshape(n, 64)
        gb.free()
        gs = mctx.testgen_scalars(args.dist, 2, n)
        pinned = cozk.PinnedBuffer(n * 32)
        pinned.array[:] = gs.download()
        gs.free()
        t0 = time.perf_counter()
        srs = mctx.srs_register(hb, sliced=True)
        rec["srs_register_ms"] = 1e3 * (time.perf_counter() - t0)
        out = np.zeros((1, 72), dtype=np.uint8)
        for _ in range(2):
            mctx.msm_batch_ptrs(srs, [pinned.ptr], n, out=out)
        wall, dev_total, dev_compute = 0.0, 0.0, 0.0
        pairs = mults = 0.0
        for _ in range(args.strong_steps):
            for d in range(n_devices):
                mctx.flush_l2(d)
            t0 = time.perf_counter()
            mctx.msm_batch_ptrs(srs, [pinned.ptr], n, out=out)
            wall += time.perf_counter() - t0
            sts = [mctx.last_stats(d) for d in range(n_devices)]
            dev_total += max(st["total_ms"] for st in sts)
            dev_compute += max(st["total_ms"] - st["h2d_ms"] for st in sts)
            pairs = sum(st["pairs"] for st in sts)
            mults = sum(st["field_mults"] for st in sts)
        k_ = args.strong_steps
        rec.update({"steps": k_, "e2e_ms": 1e3 * wall / k_, "e2e_mpoints_per_s": n / (wall / k_) / 1e6,
                    "device_ms_max": dev_total / k_, "device_mpoints_per_s": n / (dev_total / k_ * 1e-3) / 1e6,
                    "compute_ms_max": dev_compute / k_,
                    "timing_note": "e2e = wall clock around cozk_msm_batch (H2D of every device's slice, kernels, D2H, host combine); "
                                   "device_ms_max = the slowest device's CUDA-event time of its part (its H2D included); compute_ms_max "
                                   "leaves the H2D wait out",
                    "window_bits": int(sts[0]["window"]), "windows": int(sts[0]["windows"]),
                    "canonical_limb_products": n * CANON_MULTS_PER_POINT * LIMB_PRODUCTS_PER_MULT,
                    "actual_limb_products": pairs * LIMB_PRODUCTS_PER_MADD + (mults - 10.0 * pairs) * LIMB_PRODUCTS_PER_MULT,
                    "result_x_le": bytes(out[0][:8]).hex()})
        if check:
            from oracle import orc as _o
            want = _o.msm(hb, pinned.array.reshape(n, 32))
            rec["parity_vs_oracle"] = bool((want == out[0]).all())
        mctx.srs_release(srs)
        pinned.free()
    return rec


def replay_record(cozk, args, n_devices):
    """BASELINE.json configs[4] ("co-jolt 3-party proof of a 2^22-cycle guest trace, all party commitments on GPU") as a
    replay of one party's commitment-path calls (tools/replay_cojolt.py: 138 shared + 60 packed public trace polynomials,
    the single commits, 54 final_cts polynomials, the opening) through the reference-facing PST13 calls with HOST buffers,
    on one context over all GPUs of the box; next to the reference's own trace numbers.  The Rust prover cannot run here, so
    the party prove time is a projection: the trace's non-MSM time + the measured GPU time."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("replay_cojolt", os.path.join(ROOT, "tools", "replay_cojolt.py"))
    rc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rc)
    lt = args.replay_log2t
    out = {"log2_T": lt, "n_devices": n_devices, "parties": []}
    with cozk.Context(devices=list(range(n_devices))) as rctx:
        setup, srs_s = rc.make_setup(rctx, lt)
        out["srs_generate_register_s"] = round(srs_s, 3)
        for party in (0, 2):  # party 1 has the shape of party 0 (a constant share vector)
            out["parties"].append(rc.replay_party(rctx, setup, lt, party, gpus=n_devices, srs_s=srs_s))
        for h in setup.level_srs:
            rctx.srs_release(h)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-log2n", type=int, default=24, help="total points of the strong-scaling record (0 = skip)")
    ap.add_argument("--strong-steps", type=int, default=5)
    ap.add_argument("--strong-opts", default="", help="engine options for the strong record's context: name=value,name=value")
    ap.add_argument("--replay-log2t", type=int, default=22,
                    help="co-jolt party commitment-path replay for a 2^T-cycle trace (BASELINE.json configs[4]); 0 = skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    config = {"workload": "standalone BN254 G1 MSM, 2^%d %s share scalars per GPU (BASELINE.json configs[1]), "
                          "k=1, point-range sharded across GPUs" % (args.log2n, args.dist),
              "log2_points_per_gpu": args.log2n, "scalar_dist": args.dist, "scalar_form": "Fr Montgomery, stride 32",
              "l2": "flushed (256 MiB write) between timed steps", "parallelism": "point-range x%d" % world}

    if args.impl == "reference":
        if rank != 0:
            return 0
        log2n = min(args.log2n, 20 if cores >= 8 else 18)
        v, ms, sample = cpu_reference_run(log2n, args.dist, args.steps, max(args.warmup, 1), cores)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "dtype_note": "254-bit Montgomery field elements on 4 x 64-bit limbs (mulx/adc)", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "Mpoints/s", "cores": cores, "kind": "port",
                                 "sample": sample + "; C restatement of the reference's CPU Pippenger (arkworks msm_bigint_wnaf "
                                           "shape); the Rust reference itself cannot be built here"},
                "e2e": {"value": v, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    cozk = importlib.import_module("co-zkvms_b200")
    sharding = importlib.import_module("co-zkvms_b200.sharding")
    import torch
    import torch.distributed as dist
    if world > 1:
        # control plane only (barrier, max of the timings, 72-byte partial sums): gloo.  The data path has no
        # collective, so NCCL is deliberately not initialised - it would only add start-up time and
        # its version banner on stdout.
        torch.cuda.set_device(local_rank)
        dist.init_process_group("gloo", rank=rank, world_size=world)

    def barrier():
        torch.cuda.synchronize(local_rank)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    n = 1 << args.log2n
    ctx = cozk.Context(devices=[local_rank])
    if args.window:
        ctx.set_option("window", args.window)
    # rank r owns points [r*n, (r+1)*n) of the global MSM: generate only that slice, on the device
    dbases = ctx.testgen_bases(1, n, start=rank * n)
    t_reg = time.perf_counter()
    srs = ctx.srs_register_device(dbases, n)
    srs_register_ms = 1e3 * (time.perf_counter() - t_reg)
    dbases.free()
    dscal = ctx.testgen_scalars(args.dist, 2, n, start=rank * n, total_n=world * n)
    pinned = cozk.PinnedBuffer(n * 32)
    pinned.array[:] = dscal.download()

    # self-measured integer-multiply pipe peak: the rate at which the chip retires 32x32->64-bit multiply-accumulates.
    # Two estimates, the larger is the denominator: (a) carry-chained IMAD.WIDE.U32.X rows with loop-carried multipliers,
    # (b) the field multiplication itself (136 such products each).  32-bit IMAD runs at twice that rate but a product
    # needs two of them; IMAD.WIDE without carry (+ IADD3 pair) is slower (profiles/r1_microbench.md).
    sm = 148
    ms, ops = ctx.microbench("imad_cc", sm * 8, 256, 2048)
    imad_cc_rate = ops / (ms * 1e-3)
    ms_f, ops_f = ctx.microbench("fq_mul", sm * 8, 256, 512)
    fqmul_rate = ops_f / (ms_f * 1e-3)
    imad_peak = max(imad_cc_rate, fqmul_rate * LIMB_PRODUCTS_PER_MULT)
    ms_m, ops_m = ctx.microbench("madd", sm * 16, 128, 256)
    madd_rate = ops_m / (ms_m * 1e-3)

    out = np.zeros((1, 72), dtype=np.uint8)
    for _ in range(args.warmup):
        ctx.msm_batch_ptrs(srs, [dscal.ptr], n, device=0, out=out)
    # ---- device-resident timing
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    stage = {}
    dev_ms = 0.0
    launches = 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.msm_batch_ptrs(srs, [dscal.ptr], n, device=0, out=out)
        st = ctx.last_stats()
        dev_ms += st["total_ms"]
        launches += int(st["launches"])
        for k_ in ("decompose_ms", "sort_ms", "accumulate_ms", "reduce_ms", "finish_ms", "h2d_ms"):
            stage[k_] = stage.get(k_, 0.0) + st[k_]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    # ---- end to end: pinned host scalars -> result on the host
    for _ in range(2):
        ctx.msm_batch_ptrs(srs, [pinned.ptr], n, out=out)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        t0 = time.perf_counter()
        ctx.msm_batch_ptrs(srs, [pinned.ptr], n, out=out)
        e2e_s += time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    partial = out.copy()

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s * 1e3, wall_ms, srs_register_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms_total, wall_ms, srs_register_ms = [float(x) for x in t]
        e2e_s = e2e_ms_total / 1e3
        allp = sharding.gather_partials(partial)
        total_point = sharding.combine(allp, cozk.g1_sum)[0]
    else:
        total_point = partial[0]

    parity = None
    strong = None
    if rank == 0 and not args.no_cpu_baseline:
        # the checker (never the thing measured): the combined point of the timed MSM against the CPU oracle
        from oracle import orc as _o
        _o.build()
        total_n = world * n
        if total_n <= (1 << 24):
            gb = ctx.testgen_bases(1, total_n)
            hb = gb.download().reshape(total_n, 64)
            gb.free()
            want = _o.msm(hb, _o.gen_scalars(args.dist, 2, total_n))
            parity = bool((want == total_point).all())
            del hb
    if rank == 0 and args.strong_log2n:
        strong = strong_record(cozk, args, world, not args.no_cpu_baseline)
    replay = None
    if rank == 0 and args.replay_log2t:
        replay = replay_record(cozk, args, world)
    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = world * n / (ms_per_step * 1e-3) / 1e6
        e2e_value = world * n / (e2e_s / args.steps) / 1e6
        acc_ms = stage["accumulate_ms"] / args.steps
        limb_products = n * CANON_MULTS_PER_POINT * LIMB_PRODUCTS_PER_MULT
        achieved = limb_products / (acc_ms * 1e-3)
        plan_mults = st["field_mults"]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f)["k_accumulate<true>"].get("log2n=%d,dist=%s,window=%d" % (args.log2n, args.dist, int(st["window"])))
        except Exception:
            traffic = None
        line = {"metric": METRIC, "value": value, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "dtype_note": "254-bit Montgomery field elements on 8 x 32-bit limbs, "
                "IMAD.WIDE.U32 carry chains; bit-exact integer arithmetic", "data": "synthetic",
                "config": config,
                "e2e": {"value": e2e_value, "unit": "Mpoints/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 72,
                        "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": launches,
                "gpu_launches_note": "every kernel of the call is the engine's own: fused decompose + sort passes (count / scan / "
                                     "scatter), accumulate levels, reduce levels, finish",
                "roofline": {"bound": "imad", "kernel": "k_accumulate (bucket accumulation, all levels)",
                             "achieved": achieved / 1e9, "peak": imad_peak / 1e9, "unit": "G limb-products/s (IMAD.WIDE.U32 lane-ops)",
                             "frac": achieved / imad_peak, "traffic": traffic,
                             "traffic_note": "DRAM bytes per launch of k_accumulate<true> (level 1) from the committed ncu --set full capture "
                                             "(profiles/round2_accumulate_ncu.md); algorithmic bytes = pairs x 72",
                             "peak_source": "self-measured in this run: max(carry-chained IMAD.WIDE.U32.X microbenchmark, "
                                            "fq_mul microbenchmark x 136); MEASURED_PEAKS.json has no integer figure; "
                                            "nominal 148 SM x 32 lanes/clk x 1.965 GHz = 9309 G/s",
                             "imad_cc_microbench_G_per_s": imad_cc_rate / 1e9,
                             "convention": "160 field mults/point x 136 limb products (SURVEY.md 8(d)); = 43,520 IMAD lo/hi slots",
                             "kernel_ms": acc_ms,
                             "pipeline_frac": limb_products / (ms_per_step * 1e-3) / imad_peak,
                             "actual_field_mults_per_point": plan_mults / n,
                             "actual_limb_products_per_mixed_add": LIMB_PRODUCTS_PER_MADD,
                             "frac_actual_work": (st["pairs"] * LIMB_PRODUCTS_PER_MADD
                                                  + (plan_mults - 10.0 * st["pairs"]) * LIMB_PRODUCTS_PER_MULT)
                                                 / (ms_per_step * 1e-3) / imad_peak,
                             "frac_actual_work_note": "limb products the whole call really issues (dedicated squaring and the fused "
                                                      "product pair need fewer than the canonical 10 x 136 per mixed addition) / call "
                                                      "time / peak",
                             "window_bits": int(st["window"]), "windows": int(st["windows"]),
                             "fq_mul_per_s_measured": fqmul_rate, "fq_mul_frac_of_imad_peak": fqmul_rate * LIMB_PRODUCTS_PER_MULT / imad_peak,
                             "madd_per_s_measured": madd_rate, "madd_frac_of_imad_peak": madd_rate * 10 * LIMB_PRODUCTS_PER_MULT / imad_peak,
                             "hbm_algorithmic_gbs": (n * (64 + 32) + st["pairs"] * (64 + 8 * 4 * 2)) / (ms_per_step * 1e-3) / 1e9},
                "stages_ms": {k_: v / args.steps for k_, v in stage.items()},
                "stages_note": "decompose_ms = digit decomposition fused with the first sort pass; sort_ms = the remaining passes",
                "srs_register_ms": srs_register_ms,
                "wall_ms_per_step_incl_flush": wall_ms / args.steps,
                "clocks": sampler.summary(),
                "result_x_le": bytes(total_point[:8]).hex(), "result_infinity": int(total_point[64])}
        if parity is not None:
            line["parity_vs_oracle"] = parity
        if strong is not None:
            line["strong"] = strong
        if replay is not None:
            line["cojolt_replay"] = replay
        if not args.no_cpu_baseline and world == 1:
            log2c = min(args.log2n, 20 if cores >= 8 else 18)
            v, ms_c, sample = cpu_reference_run(log2c, args.dist, 2, 1, cores)
            anchor = 0.125 * cores
            line["cpu_baseline"] = {"value": v, "unit": "Mpoints/s", "cores": cores, "kind": "port", "sample": sample,
                                    "ms_per_msm": ms_c, "per_core": v / cores, "vs_anchor": v / anchor,
                                    "anchor": "the reference's own traces give 0.12-0.13 Mpoints/s per vCPU for arkworks / jolt-core "
                                              "(SURVEY.md section 6); vs_anchor = this port / (0.125 Mpoints/s x cores): below 1 the "
                                              "port is slower than the real reference and every GPU/CPU ratio is inflated by 1 / vs_anchor"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
