"""GPU tier: the batched-affine pre-reduction (csrc/affine_kernels.cuh) through the test entry point of the C ABI
(include/cozk_test.h) - the engine's batched kernels (shared inversions, product trees) against the serial contract kernel
on the GPU and against the contract body run on the host (tests/emul) - and whole MSMs with the rounds switched on against
the CPU oracle."""
import importlib

import numpy as np
import pytest

from tests import emul

pytestmark = pytest.mark.gpu
SKIP, NEG = 0xFFFFFFFF, 0x80000000


@pytest.fixture(scope="module")
def ctx():
    cozk = importlib.import_module("co-zkvms_b200")
    with cozk.Context() as c:
        yield c


def _random_pairs(rng, buckets, max_run, npts, p_skip=0.05, dup=True):
    keys, vals = [], []
    for b in range(buckets):
        ln = int(rng.integers(0, max_run + 1))
        keys += [b * 3 + 1] * ln
        v = rng.integers(0, npts, size=ln, dtype=np.uint64).astype(np.uint32)
        if dup and ln >= 4:
            v[1] = v[0]                                   # P + P
            v[3] = v[2]
        v = v | np.where(rng.random(ln) < 0.5, NEG, 0).astype(np.uint32)
        if dup and ln >= 4:
            v[3] = v[2] ^ np.uint32(NEG)                  # P + (-P)
        v[rng.random(ln) < p_skip] = SKIP
        vals += v.tolist()
    return np.array(keys, np.uint32), np.array(vals, np.uint32)


def _same_lists(a, b):
    (ka, va, pa), oa = a
    (kb, vb, pb), ob = b
    assert (ka == kb).all() and (va == vb).all()
    live = va != SKIP
    assert (pa[live] == pb[live]).all()
    assert len(oa) == len(ob)
    for (k1, p1), (k2, p2) in zip(oa, ob):
        o1, o2 = np.argsort(k1, kind="stable"), np.argsort(k2, kind="stable")
        assert (k1[o1] == k2[o2]).all() and len(set(k1.tolist())) == len(k1)
        assert (p1[o1] == p2[o2]).all()


@pytest.mark.parametrize("buckets,max_run,rounds", [(1, 1, 1), (1, 5, 3), (3, 2, 2), (40, 9, 3), (700, 30, 3), (64, 4000, 4), (5000, 3, 2),
                                                      (9, 70000, 6)])
def test_batched_rounds_match_contract_kernel(ctx, orc, buckets, max_run, rounds):
    rng = np.random.default_rng(buckets * 131 + max_run)
    npts = 5000
    dpts = ctx.testgen_bases(3, npts)
    keys, vals = _random_pairs(rng, buckets, max_run, npts)
    if keys.size < 2:
        keys, vals = np.array([4, 4], np.uint32), np.array([1, 2], np.uint32)
    got = ctx.affine_rounds(keys, vals, dpts, 3 * buckets + 2, rounds)
    ref = ctx.affine_rounds(keys, vals, dpts, 3 * buckets + 2, rounds, reference=True)
    _same_lists(got[:2], ref[:2])
    dpts.free()


def test_batched_round_matches_host_contract(ctx, orc):
    """One round on the GPU against the contract body compiled for the host (the body the CPU tier checks against Python
    big-integer arithmetic)."""
    rng = np.random.default_rng(77)
    npts = 600
    bases = orc.gen_bases(3, npts)
    dpts = ctx.testgen_bases(3, npts)
    keys, vals = _random_pairs(rng, 300, 12, npts)
    (gk, gv, gp), govf, _ = ctx.affine_rounds(keys, vals, dpts, 1000, 1)
    (hk, hv, hp), (hok, hop) = emul.affine_round(keys, vals, bases)
    _same_lists(((gk, gv, gp), govf), ((hk, hv, hp), [(hok, hop)]))
    dpts.free()


@pytest.mark.parametrize("dist", ["uniform", "const", "wminus", "dup", "zero_half"])
@pytest.mark.parametrize("log2n", [14, 18])
def test_msm_with_rounds_matches_oracle(ctx, orc, dist, log2n):
    """Whole MSMs with the pre-reduction on (table mode and per-window buckets, one vector and a batch, a chunked host call)."""
    n = 1 << log2n
    bases = orc.gen_bases(1, n)
    srs = ctx.srs_register(bases)
    v = orc.gen_scalars(dist, 9, n)
    want = orc.msm(bases, v)
    try:
        ctx.set_option("affine_min_pairs", 0)
        for rounds in (1, 3, 5):
            ctx.set_option("affine_rounds", rounds)
            assert (ctx.msm_batch(srs, [v], n=n)[0] == want).all(), (dist, rounds)
        ctx.set_option("affine_rounds", 3)
        if log2n == 14:
            ctx.set_option("window", 9)   # forced window: per-window bucket sets, no table
            assert (ctx.msm_batch(srs, [v], n=n)[0] == want).all()
            ctx.set_option("window", 0)
            got = ctx.msm_batch(srs, [v, v, v], n=n)
            assert all((g == want).all() for g in got)
        else:
            ctx.set_option("stream_min_points", 1 << 16)
            ctx.set_option("stream_chunks", 3)
            assert (ctx.msm_batch(srs, [v], n=n)[0] == want).all()
    finally:
        for k, val in (("affine_rounds", 0), ("affine_min_pairs", 1 << 20), ("window", 0), ("stream_min_points", 1 << 20), ("stream_chunks", 0)):
            ctx.set_option(k, val)
        ctx.srs_release(srs)
