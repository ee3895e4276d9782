"""CPU tier: the batched-affine pre-reduction (co-zkvms_b200/csrc/affine_kernels.cuh).  Its contract body - one halving round,
out[o] = in[2o] + in[2o + 1] inside a bucket, the first point of a bucket that starts at an odd position to the overflow
list - against Python big-integer affine arithmetic (oracle/pyref.py), every exceptional case included; then the whole
pipeline with the rounds in front of the accumulate levels against the oracle MSM."""
import numpy as np
import pytest

from oracle import pyref
from tests import emul
from tests import helpers as H

SKIP, NEG = 0xFFFFFFFF, 0x80000000


def _pt(row):
    return (pyref.from_mont(H.to_int(row[:32]), H.P), pyref.from_mont(H.to_int(row[32:64]), H.P))


def _want_round(keys, vals, pts):
    """Python restatement: per output the expected (key, point or None), and the overflow dict key -> point."""
    def load(v):
        if v == SKIP:
            return None
        p = pts[v & ~NEG]
        return pyref.neg(p) if v & NEG else p
    out, ovf = [], {}
    for o in range((len(keys) + 1) // 2):
        k0, p0 = keys[2 * o], load(vals[2 * o])
        if 2 * o + 1 >= len(keys):
            out.append((k0, p0))
            continue
        k1, p1 = keys[2 * o + 1], load(vals[2 * o + 1])
        if k1 != k0:
            out.append((k0, p0))
            if p1 is not None:
                assert k1 not in ovf
                ovf[k1] = p1
        else:
            out.append((k0, pyref.add(p0, p1)))
    return out, ovf


def _check_round(keys, vals, pts_int):
    wire = H.bases_wire(pts_int)
    (ko, vo, po), (ok, op) = emul.affine_round(np.array(keys, np.uint32), np.array(vals, np.uint32), wire)
    want, wovf = _want_round(keys, vals, pts_int)
    assert len(ko) == len(want)
    for o, (k, p) in enumerate(want):
        assert ko[o] == k
        if p is None:
            assert vo[o] == SKIP, o
        else:
            assert vo[o] == o and _pt(po[o]) == p, o
    assert sorted(ok.tolist()) == sorted(wovf), "overflow keys"
    for k, row in zip(ok.tolist(), op):
        assert _pt(row) == wovf[k]
    return (ko, vo, po), (ok, op)


def test_round_body_against_python():
    pts = [pyref.base_point(3, i) for i in range(12)]
    pts[5] = pts[4]                      # a duplicated base: P + P is a doubling
    keys = [0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 2, 3, 5, 5, 5, 5, 7, 7, 9, 9, 9, 9]
    vals = [0, 1, 2 | NEG, 3,            # plain additions, one negated operand
            4, 5, 6,                     # doubling (4 == 5), then a run that ends at an even position
            7, 7 | NEG, SKIP, 8, SKIP,   # (7, 2: new run at odd position -> overflow) ... P + (-P) = null, skip + P, lone skip
            9,                           # single-entry bucket
            10, 10, 11, SKIP,            # P + P through the same index
            0, SKIP, 1, 1 | NEG, SKIP, SKIP]
    assert len(keys) == len(vals)
    _check_round(keys, vals, pts)
    # odd length: the last entry passes through alone; empty overflow when every run starts at an even position
    _check_round([4, 4, 4, 4, 6, 6, 8], [0, 1, 2, 3, 4, 6, 7], pts)
    _check_round([1], [3 | NEG], pts)


def test_rounds_compose_to_bucket_sums():
    """Three rounds over random runs, then every bucket's entries (reduced list + overflow lists) sum to the bucket's points."""
    rng = np.random.default_rng(5)
    npts = 40
    pts = [pyref.base_point(9, i) for i in range(npts)]
    keys, vals = [], []
    for b in range(14):
        for _ in range(int(rng.integers(0, 12))):
            keys.append(b)
            v = int(rng.integers(0, npts))
            vals.append(SKIP if rng.random() < 0.1 else (v | (NEG if rng.random() < 0.5 else 0)))
    want = {}
    for k, v in zip(keys, vals):
        if v != SKIP:
            p = pts[v & ~NEG]
            want[k] = pyref.add(want.get(k), pyref.neg(p) if v & NEG else p)
    got = {}
    k_in, v_in, p_int = keys, vals, pts
    for _ in range(3):
        (ko, vo, po), (ok, op) = _check_round(k_in, v_in, p_int)
        for k, row in zip(ok.tolist(), op):
            got[k] = pyref.add(got.get(k), _pt(row))
        k_in, v_in = ko.tolist(), vo.tolist()
        p_int = [_pt(r) if v != SKIP else None for r, v in zip(po, v_in)]
        p_int = [p if p is not None else pts[0] for p in p_int]  # never read: their vals carry the skip mark
    for k, v in zip(k_in, v_in):
        if v != SKIP:
            got[k] = pyref.add(got.get(k), p_int[v])
    for k in set(want) | set(got):
        assert got.get(k) == want.get(k), k


@pytest.mark.parametrize("dist", pyref.DISTS)
def test_msm_with_affine_rounds(orc, dist):
    """The whole pipeline with 1 .. 4 rounds in front of the accumulate levels, runs much longer and much shorter than 2^rounds,
    plain and table mode, chunked calls: the same point as the oracle's."""
    n = 700
    bases = orc.gen_bases(11, n)
    sc = orc.gen_scalars(dist, 6, n)
    want = orc.msm(bases, sc)
    try:
        for rounds in (-1, -2, -3, -4):
            emul.set_affine_rounds(rounds)
            for kw in ({"c": 3}, {"c": 7}, {"table_c": 5}, {"table_c": 9, "stream_chunks": 3}):
                got, _ = emul.msm(bases, sc, **kw)
                assert (got[0] == want).all(), (dist, rounds, kw)
    finally:
        emul.set_affine_rounds(0)


def test_round_rule_and_batches(orc):
    """The engine's rule (rounds only while the average bucket keeps 4 entries) and a batch with several vectors, strides and
    points at infinity."""
    n, g = 900, 3
    bases = orc.gen_bases(2, n)
    vecs = [orc.gen_scalars(d, 30 + i, n, stride=64) for i, d in enumerate(("uniform", "const", "zero_half"))]
    inf = np.zeros(n, np.uint8)
    inf[5::7] = 1
    masked = [v.copy() for v in vecs]
    for v in masked:
        v[5::7] = 0
    try:
        emul.set_affine_rounds(3)
        got, st = emul.msm(bases, np.concatenate(vecs), g=g, stride=64, infinity=inf, table_c=4)  # 8 buckets per vector: long runs
        reduced, overflow = emul.affine_stats()
        assert reduced == (int(st[4]) + 7) // 8 and overflow > 0
        for j in range(g):
            assert (got[j] == orc.msm(bases, masked[j])).all(), j
        got, st = emul.msm(bases, vecs[0], stride=64, c=12)  # 2048 buckets per window: runs too short, no round
        assert emul.affine_stats()[0] == int(st[4]) and (got[0] == orc.msm(bases, vecs[0])).all()
    finally:
        emul.set_affine_rounds(0)
