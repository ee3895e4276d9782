"""CPU tier: the N>1 path's host logic (shard ranges, gloo gather of partial sums, host-side combine) with
world_size 2 on the gloo backend.  The per-rank partial sums come from the oracle here (no GPU); on the GPU box the
same functions carry the engine's partial sums (bench.py --gpus N)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    sh = importlib.import_module("co-zkvms_b200.sharding")
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            rs = [sh.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import orc
    cozk = importlib.import_module("co-zkvms_b200")
    sh = importlib.import_module("co-zkvms_b200.sharding")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n, k = 3001, 2
    bases = orc.gen_bases(1, n)
    vecs = [orc.gen_scalars(d, 9, n) for d in ("uniform", "const")]
    lo, hi = sh.shard_range(n, rank, world)
    partial = np.stack([orc.msm(bases[lo:hi], v[lo:hi], threads=2) for v in vecs])
    allp = sh.gather_partials(partial)
    total = sh.combine(allp, cozk.g1_sum)
    want = np.stack([orc.msm(bases, v, threads=2) for v in vecs])
    q.put((rank, bool((total == want).all()), allp.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_combine():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] == (2, 2, 72) for r in res)


def test_spartan_mirror_eta_powers():
    """co-zkvms_b200/spartan.py: the `x *= eta` sequence of aggregate_poly (co-spartan/src/utils.rs:96-102) as Montgomery
    coefficients for the device-side linear combination."""
    from oracle import pyref
    from tests import helpers as H
    sp = importlib.import_module("co-zkvms_b200.spartan")
    eta = pyref.scalar_uniform(5, 0)
    got = sp._fr_powers(H.fr_mont(eta), 6)
    assert [pyref.from_mont(H.to_int(row), H.R) for row in got] == [pow(eta, j, H.R) for j in range(6)]
    assert sp._fr_powers(H.fr_mont(0), 3).tolist() == [list(H.fr_mont(1)), list(H.fr_mont(0)), list(H.fr_mont(0))]


def test_bench_reference_arm_line():
    """bench.py --impl reference runs the CPU restatement only (no GPU, no engine) and prints the contract's JSON line."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2n", "14", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpoints/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["config"]["log2_points_per_gpu"] == 14


def test_committed_bench_line_has_the_contract_keys():
    """The last bench line measured on the B200 (profiles/round2_bench_1gpu.json, copied from the gpurun call) carries every key
    the driver's contract names; a change of bench.py that drops one shows up here as soon as the line is refreshed."""
    import json
    with open(os.path.join(ROOT, "profiles", "round2_bench_1gpu.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "parity_vs_oracle", "strong",
              "cojolt_replay", "srs_register_ms"):
        assert k in line, k
    assert line["config"]["workload"] and "model" not in line["config"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] == 32 << line["config"]["log2_points_per_gpu"] and line["e2e"]["value"] < line["value"]
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "port"
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and line["clocks"]["samples"] >= 10
    assert line["gpu_launches"] > 0 and line["parity_vs_oracle"] is True and line["warmup"] >= 3
