"""Shared helpers of the test tiers (wire-format conversions, golden fixture access)."""
import json
import os

import numpy as np

from oracle import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
P, R = pyref.P, pyref.R_ORDER


def golden():
    with open(os.path.join(HERE, "golden", "bn254_msm_golden.json")) as f:
        return json.load(f)


def le32(x):
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint8)


def to_int(b):
    return int.from_bytes(bytes(bytearray(b)), "little")


def fq_mont(x):
    return le32(x * pyref.MONT_R % P)


def fr_mont(x):
    return le32(x * pyref.MONT_R % R)


def scalars_wire(values, form=0, stride=32):
    """canonical ints -> (n, stride) uint8, Montgomery (form 0) or canonical (form 1)."""
    out = np.zeros((len(values), stride), dtype=np.uint8)
    for i, v in enumerate(values):
        out[i, :32] = fr_mont(v) if form == 0 else le32(v)
    return out


def bases_wire(points):
    out = np.zeros((len(points), 64), dtype=np.uint8)
    for i, (x, y) in enumerate(points):
        out[i, :32] = fq_mont(x)
        out[i, 32:] = fq_mont(y)
    return out


def point_wire(pt):
    out = np.zeros(72, dtype=np.uint8)
    if pt is None:
        out[64] = 1
    else:
        out[:32] = fq_mont(pt[0])
        out[32:64] = fq_mont(pt[1])
    return out


def parse_point(js):
    return None if js is None else (int(js[0], 16), int(js[1], 16))


def golden_msm_inputs(case):
    """(bases (n,64), scalars canonical ints) of one golden MSM case."""
    n = case["n"]
    pts = [pyref.base_point(case["base_seed"], i) for i in range(n)]
    if case.get("dup_base0"):
        pts[1] = pts[0]
    if case["dist"].startswith("explicit:"):
        sc = [int(s, 16) for s in case["scalars"]]
    else:
        sc = pyref.scalars(case["dist"], case["scalar_seed"], n)
    return pts, sc


def multi_device_ids(want=4):
    """Device ids for a multi-device context: the visible GPUs (at most `want`), or - on a single-GPU box - the SAME GPU
    three times: the engine then drives three logical devices (own streams, scratch and SRS copies / slices) on one
    physical GPU, which exercises every host-side sharding path (threads, slices, partial sums, peer-copy fallbacks)."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev >= 2:
        return list(range(min(ndev, want)))
    return [0, 0, 0]
