"""GPU tier: the hand-written pair sort (csrc/sort_kernels.cuh) on its own, through the test entry point of the C ABI
(include/cozk_test.h).  The sort groups pairs by key and leaves the order inside a group open (it does not matter to a
sum), so outputs are compared as: keys ascending, and per key the same multiset of vals as the input.  The fused first
pass (pairs produced from the scalars inside the sort) is compared with the decompose kernel's pairs the same way."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    cozk = importlib.import_module("co-zkvms_b200")
    with cozk.Context() as c:
        yield c


def _check_sorted(keys, vals, key_bits, got_k, got_v):
    assert int(keys.max()) < (1 << key_bits)
    assert (np.diff(got_k.astype(np.int64)) >= 0).all(), "keys not ascending"
    order = np.lexsort((vals, keys))
    gorder = np.lexsort((got_v, got_k))
    assert (got_k[gorder] == keys[order]).all()
    assert (got_v[gorder] == vals[order]).all()


@pytest.mark.parametrize("m", [1, 2, 31, 32, 33, 511, 512, 513, 8191, 8192, 8193, 100_000, (1 << 20) + 12345])
@pytest.mark.parametrize("key_bits", [1, 5, 8, 9, 16, 17, 21, 24, 25, 31])
def test_generic_passes_group_by_key(ctx, m, key_bits):
    if m > (1 << 17) and key_bits not in (16, 21, 31):
        pytest.skip("large sizes on the key widths the engine uses most")
    rng = np.random.default_rng(m * 37 + key_bits)
    keys = rng.integers(0, 1 << key_bits, size=m, dtype=np.uint64).astype(np.uint32)
    vals = rng.integers(0, 1 << 32, size=m, dtype=np.uint64).astype(np.uint32)
    got_k, got_v = ctx.sort_pairs(keys, vals, key_bits)
    _check_sorted(keys, vals, key_bits, got_k, got_v)


@pytest.mark.parametrize("digit_bits", [7, 9, 10, 11])
def test_other_digit_widths(ctx, digit_bits):
    """Option sort_digit_bits: passes of 7 .. 11 bits (11 = the whole shared-memory counter window of a tile; rows of up to
    2048 entries in the row scan)."""
    ctx.set_option("sort_digit_bits", digit_bits)
    try:
        for m, key_bits in ((70_000, 21), (200_000, 16), (5_000, 25), (40_000, 11)):
            rng = np.random.default_rng(m + digit_bits)
            keys = rng.integers(0, 1 << key_bits, size=m, dtype=np.uint64).astype(np.uint32)
            vals = np.arange(m, dtype=np.uint32)
            got_k, got_v = ctx.sort_pairs(keys, vals, key_bits)
            _check_sorted(keys, vals, key_bits, got_k, got_v)
        dsc = ctx.testgen_scalars("uniform", 5, 30_000)
        plain_k, plain_v = ctx.decompose_sort(dsc, 30_000, 20, table_stride=30_000, key_bits=0, fused=False)
        got_k, got_v = ctx.decompose_sort(dsc, 30_000, 20, table_stride=30_000, key_bits=19, fused=True)
        wk, wv = _canon(plain_k, plain_v)
        gk, gv = _canon(got_k, got_v)
        assert (np.diff(got_k.astype(np.int64)) >= 0).all() and (gk == wk).all() and (gv == wv).all()
        dsc.free()
    finally:
        ctx.set_option("sort_digit_bits", 8)


@pytest.mark.parametrize("shape", ["all_equal", "two_values", "sorted", "reversed", "few_hot"])
def test_degenerate_key_distributions(ctx, shape):
    """Constant share vectors put every pair of a window into ONE bucket: the sort must not care."""
    m, key_bits = 300_000, 16
    rng = np.random.default_rng(5)
    if shape == "all_equal":
        keys = np.full(m, 0xBEEF, np.uint32)
    elif shape == "two_values":
        keys = np.where(rng.integers(0, 2, m) == 1, 0xFFFF, 0).astype(np.uint32)
    elif shape == "sorted":
        keys = np.sort(rng.integers(0, 1 << key_bits, m)).astype(np.uint32)
    elif shape == "reversed":
        keys = np.sort(rng.integers(0, 1 << key_bits, m))[::-1].astype(np.uint32)
    else:
        keys = rng.integers(0, 1 << key_bits, m).astype(np.uint32)
        keys[rng.integers(0, m, m // 2)] = 77
    vals = np.arange(m, dtype=np.uint32)
    got_k, got_v = ctx.sort_pairs(keys, vals, key_bits)
    _check_sorted(keys, vals, key_bits, got_k, got_v)


@pytest.mark.parametrize("key_bits,hot", [(16, 3), (21, 1), (24, 40)])
def test_tiny_partitions_take_the_slow_path(ctx, key_bits, hot):
    """Few pairs spread over many key values: the partitions of the earlier passes are far smaller than a tile, so most
    pairs fall outside the tile's shared-memory window and take their own global slot.  Plus a few huge buckets."""
    m = 60_000
    rng = np.random.default_rng(key_bits)
    keys = rng.integers(0, 1 << key_bits, size=m, dtype=np.uint64).astype(np.uint32)
    for h in range(hot):
        keys[rng.integers(0, m, m // (2 * hot))] = rng.integers(0, 1 << key_bits)
    vals = np.arange(m, dtype=np.uint32)
    got_k, got_v = ctx.sort_pairs(keys, vals, key_bits)
    _check_sorted(keys, vals, key_bits, got_k, got_v)


def _canon(keys, vals):
    order = np.lexsort((vals, keys))
    return keys[order], vals[order]


@pytest.mark.parametrize("n,c,g,table", [(1, 16, 1, False), (37, 3, 1, False), (1000, 8, 3, False), (4096, 11, 2, False),
                                         (100_003, 16, 1, False), (1 << 17, 17, 1, True), (5000, 13, 5, True),
                                         (70_000, 20, 1, True), (300, 22, 1, False)])
@pytest.mark.parametrize("dist", ["uniform", "const", "small16"])
def test_fused_first_pass_equals_decompose_then_sort(ctx, n, c, g, table, dist):
    W = (255 + c - 1) // c
    B = 1 << (c - 1)
    buckets = g * (1 if table else W) * B
    key_bits = max(1, int(np.ceil(np.log2(buckets))))
    if key_bits > 25:
        pytest.skip("more buckets than the engine ever plans")
    vstride = ((n - 1) * 32 + 32 + 255) & ~255
    dsc = ctx.alloc(g * vstride)
    try:
        for v in range(g):
            one = ctx.testgen_scalars(dist, 11 + v, n)
            dsc.upload(one.download(n * 32), offset=v * vstride)
            one.free()
        tstride = n + 7 if table else 0
        plain_k, plain_v = ctx.decompose_sort(dsc, n, c, g=g, table_stride=tstride, val_offset=3 if table else 0, key_bits=0,
                                              fused=False)
        assert int(plain_k.max()) < buckets
        for fused in (True, False):
            got_k, got_v = ctx.decompose_sort(dsc, n, c, g=g, table_stride=tstride, val_offset=3 if table else 0,
                                              key_bits=key_bits, fused=fused)
            assert (np.diff(got_k.astype(np.int64)) >= 0).all(), "not sorted"
            wk, wv = _canon(plain_k, plain_v)
            gk, gv = _canon(got_k, got_v)
            assert (gk == wk).all() and (gv == wv).all(), "pairs differ (fused=%s)" % fused
    finally:
        dsc.free()
