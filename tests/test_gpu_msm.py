"""GPU tier: the MSM through the C ABI (cozk_msm_batch*), bit-exact against the oracle and the golden fixtures."""
import numpy as np
import pytest

from oracle import pyref
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_golden_vectors(ctx):
    for case in H.golden()["msm"]:
        pts, sc = H.golden_msm_inputs(case)
        want = H.point_wire(H.parse_point(case["result"]))
        srs = ctx.srs_register(H.bases_wire(pts))
        for form in (0, 1):
            got = ctx.msm_batch(srs, H.scalars_wire(sc, form), form=form)
            assert (got[0] == want).all(), (case["n"], case["dist"], form)
        ctx.srs_release(srs)


@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 33, 100, 1000, 4097, 1 << 14])
def test_sizes_and_distributions(ctx, orc, n):
    bases = orc.gen_bases(1, n)
    srs = ctx.srs_register(bases)
    for dist in pyref.DISTS:
        sc = orc.gen_scalars(dist, 3, n)
        got = ctx.msm_batch(srs, sc)
        assert (got[0] == orc.msm(bases, sc)).all(), (n, dist)
    ctx.srs_release(srs)


def test_every_window_size(ctx, orc):
    n = 3000
    bases = orc.gen_bases(2, n)
    srs = ctx.srs_register(bases)
    u = orc.gen_scalars("uniform", 4, n)
    c0 = orc.gen_scalars("const", 4, n)
    wu, wc = orc.msm(bases, u), orc.msm(bases, c0)
    try:
        for c in range(2, 23):
            ctx.set_option("window", c)
            assert (ctx.msm_batch(srs, u)[0] == wu).all(), c
            assert (ctx.msm_batch(srs, c0)[0] == wc).all(), c
    finally:
        ctx.set_option("window", 0)
    ctx.srs_release(srs)


def test_2pow18_all_party_distributions(ctx, orc):
    """The three co-jolt party share shapes and the co-spartan one at 2^18."""
    n = 1 << 18
    bases = orc.gen_bases(1, n)
    srs = ctx.srs_register(bases)
    for dist in ("uniform", "const", "wminus", "dup"):
        sc = orc.gen_scalars(dist, 2, n)
        assert (ctx.msm_batch(srs, sc)[0] == orc.msm(bases, sc)).all(), dist
    ctx.srs_release(srs)


def test_batch_strides_forms_prefix_offset(ctx, orc):
    n_srs, n = 5000, 2048
    bases = orc.gen_bases(3, n_srs)
    srs = ctx.srs_register(bases)
    assert ctx.srs_len(srs) == n_srs
    # a batch of vectors against the prefix bases[..n] (pst13.rs:319-323), Rep3 AoS stride
    dists = ["uniform", "const", "wminus", "dup", "zero_half", "small16", "uniform"]
    vecs = [orc.gen_scalars(d, 30 + i, n, stride=64) for i, d in enumerate(dists)]
    got = ctx.msm_batch(srs, vecs, n=n, stride=64)
    for j, v in enumerate(vecs):
        assert (got[j] == orc.msm(bases[:n], v)).all(), j
    # canonical form, a split_ck-style slice in the middle of the SRS (co-spartan/src/utils.rs:38-83)
    can = orc.gen_scalars("uniform", 8, n, form=1)
    got = ctx.msm_batch(srs, can, n=n, base_offset=1500, form=1)
    assert (got[0] == orc.msm(bases[1500:1500 + n], can, form=1)).all()
    # small groups: force several vector groups through the double-buffered staging
    ctx.set_option("group_pairs", 40000)
    try:
        got = ctx.msm_batch(srs, vecs, n=n, stride=64)
        for j, v in enumerate(vecs):
            assert (got[j] == orc.msm(bases[:n], v)).all(), j
    finally:
        ctx.set_option("group_pairs", 1 << 29)
    # max_num_bits hint (public u16 polynomials)
    small = orc.gen_scalars("small16", 5, n)
    assert (ctx.msm_batch(srs, small, n=n, max_num_bits=16)[0] == orc.msm(bases[:n], small)).all()
    flags = np.zeros((n, 32), np.uint8)
    flags[::3, 0] = 1
    assert (ctx.msm_batch(srs, flags, n=n, form=1, max_num_bits=1)[0] == orc.msm(bases[:n], flags, form=1)).all()
    ctx.srs_release(srs)


def test_edge_cases_and_errors(cozk, ctx, orc):
    n = 64
    bases = orc.gen_bases(4, n)
    srs = ctx.srs_register(bases)
    sc = orc.gen_scalars("uniform", 1, n)
    # empty input and all-zero scalars give the identity
    assert ctx.msm_batch(srs, [sc], n=0)[0][64] == 1
    assert ctx.msm_batch(srs, np.zeros((n, 32), np.uint8))[0][64] == 1
    # canonical scalars >= r are integers, i.e. reduced mod r
    big = np.full((n, 32), 0xFF, np.uint8)
    red = np.tile(H.le32(((1 << 256) - 1) % H.R), (n, 1))
    assert (ctx.msm_batch(srs, big, form=1)[0] == orc.msm(bases, red, form=1)).all()
    # key length error (pst13.rs:311-316), bad handle, bad stride
    with pytest.raises(cozk.CozkError) as e:
        ctx.msm_batch(srs, np.zeros((n + 1, 32), np.uint8))
    assert e.value.code == cozk.ERR_KEY_LENGTH
    with pytest.raises(cozk.CozkError) as e:
        ctx.msm_batch(srs, sc, n=8, base_offset=n - 4)
    assert e.value.code == cozk.ERR_KEY_LENGTH
    with pytest.raises(cozk.CozkError) as e:
        ctx.msm_batch(srs + 1000, sc)
    assert e.value.code == cozk.ERR_BAD_HANDLE
    with pytest.raises(cozk.CozkError) as e:
        ctx.msm_batch(srs, sc, n=4, stride=24)
    assert e.value.code == cozk.ERR_INVALID_ARG
    ctx.srs_release(srs)
    # infinity flags and the 72-byte arkworks Affine stride
    b72 = np.zeros((n, 72), np.uint8)
    b72[:, :64] = bases
    inf = np.zeros(n, np.uint8)
    inf[5::9] = 1
    srs = ctx.srs_register(b72, infinity=inf)
    masked = sc.copy()
    masked[5::9] = 0
    assert (ctx.msm_batch(srs, sc)[0] == orc.msm(bases, masked)).all()
    ctx.srs_release(srs)
    # duplicated bases: P + P inside a bucket; opposite scalars: P + (-P)
    dup = bases.copy()
    dup[1::2] = dup[0::2]
    srs = ctx.srs_register(dup)
    c0 = orc.gen_scalars("const", 2, n)
    assert (ctx.msm_batch(srs, c0)[0] == orc.msm(dup, c0)).all()
    pm = np.tile(H.le32(H.R - 1), (n, 1))
    pm[0::2] = H.le32(1)
    assert ctx.msm_batch(srs, pm, form=1)[0][64] == 1
    # a batch of more than four vectors is finished on the device and normalised on the host with ONE inversion: identity
    # results (first, in the middle, last) must not disturb their neighbours
    u = [orc.gen_scalars("uniform", 90 + j, n) for j in range(4)]
    canon = lambda v: np.stack([H.le32(pyref.from_mont(H.to_int(r), H.R)) for r in v])  # noqa: E731
    zero = np.zeros((n, 32), np.uint8)
    batch = [pm, canon(u[0]), canon(u[1]), zero, canon(u[2]), pm, canon(u[3]), zero]
    got = ctx.msm_batch(srs, batch, n=n, form=1)
    for j, v in enumerate(batch):
        assert (got[j] == orc.msm(dup, v, form=1)).all(), j
    assert got[0][64] == 1 and got[3][64] == 1 and got[5][64] == 1 and got[7][64] == 1 and got[1][64] == 0
    ctx.srs_release(srs)


def test_device_resident_scalars(ctx, orc):
    n = 1 << 15
    dbases = ctx.testgen_bases(1, n)
    srs = ctx.srs_register_device(dbases, n)
    ds = [ctx.testgen_scalars(d, 6, n, stride=64) for d in ("uniform", "const")]
    got = ctx.msm_batch_ptrs(srs, [d.ptr for d in ds], n, stride=64, device=0)
    bases = orc.gen_bases(1, n)
    for j, d in enumerate(("uniform", "const")):
        assert (got[j] == orc.msm(bases, orc.gen_scalars(d, 6, n))).all(), d
    st = ctx.last_stats()
    assert st["launches"] >= 5 and st["pairs"] > 0
    for d in ds:
        d.free()
    dbases.free()
    ctx.srs_release(srs)


@pytest.mark.parametrize("log2n", [22, 24])
def test_properties_at_full_size(ctx, log2n):
    """Size-independent checks at sizes the oracle is not run at (2^22: BASELINE.json configs[2] and [4]; 2^24: the low
    end of configs[3] on one GPU): split = sum of parts, window-size independence, constant vector = c * sum(P_i), all
    computed on the device and compared bit for bit."""
    import importlib
    cozk = importlib.import_module("co-zkvms_b200")
    n = 1 << log2n
    dbases = ctx.testgen_bases(1, n)
    srs = ctx.srs_register_device(dbases, n)
    s = ctx.testgen_scalars("uniform", 2, n, form=1)
    full = ctx.msm_batch_ptrs(srs, [s.ptr], n, form=1, device=0)[0]
    # split at an odd boundary
    cut = 1234567 if log2n == 22 else 9876543
    a = ctx.msm_batch_ptrs(srs, [s.ptr], cut, form=1, device=0)[0]
    b = ctx.msm_batch_ptrs(srs, [s.ptr + 32 * cut], n - cut, base_offset=cut, form=1, device=0)[0]
    assert (cozk.g1_sum(np.stack([a, b])) == full).all()
    # different window sizes give the same point
    ctx.set_option("window", 13)
    try:
        assert (ctx.msm_batch_ptrs(srs, [s.ptr], n, form=1, device=0)[0] == full).all()
    finally:
        ctx.set_option("window", 0)
    # constant vector: MSM(c, P) = c * MSM(1, P); check through c = 2: MSM(2) = MSM(1) + MSM(1)
    ones = np.zeros((n, 32), np.uint8)
    ones[:, 0] = 1
    twos = ones.copy()
    twos[:, 0] = 2
    d1 = ctx.alloc(n * 32).upload(ones)
    d2 = ctx.alloc(n * 32).upload(twos)
    p1 = ctx.msm_batch_ptrs(srs, [d1.ptr], n, form=1, device=0, max_num_bits=2)[0]
    p2 = ctx.msm_batch_ptrs(srs, [d2.ptr], n, form=1, device=0)[0]
    assert (cozk.g1_sum(np.stack([p1, p1])) == p2).all()
    for d in (s, d1, d2, dbases):
        d.free()
    ctx.srs_release(srs)


def test_multi_device_context(cozk, orc):
    """One process driving several GPUs (the Rust caller's shape): k >= devices shards by vector, otherwise by point
    range with the partial sums added on the host (split_ck + combine_comm).  On a single-GPU box the context is opened
    over the same GPU three times (helpers.multi_device_ids), so the sharding paths always run."""
    ids = H.multi_device_ids(8)
    ndev = len(ids)
    n = 1 << 15
    bases = orc.gen_bases(1, n)
    with cozk.Context(devices=ids) as mctx:
        assert mctx.device_count == ndev
        srs = mctx.srs_register(bases)
        dists = ["uniform", "const", "wminus", "dup", "zero_half"] * 2
        vecs = [orc.gen_scalars(d, 50 + i, n, stride=64) for i, d in enumerate(dists[: ndev + 1])]
        got = mctx.msm_batch(srs, vecs, n=n, stride=64)            # by vector
        for j, v in enumerate(vecs):
            assert (got[j] == orc.msm(bases, v)).all(), j
        one = mctx.msm_batch(srs, vecs[:1], n=n, stride=64)        # by point range
        assert (one[0] == orc.msm(bases, vecs[0])).all()
        odd = mctx.msm_batch(srs, vecs[:1], n=1001, base_offset=17, stride=64)
        assert (odd[0] == orc.msm(bases[17:1018], vecs[0][:1001])).all()
        mctx.srs_release(srs)


def test_sliced_srs(cozk, orc):
    """cozk_srs_register_sliced: every device keeps only its point range and its own table (the reference's
    split_ck, co-noir-spartan/co-spartan/src/utils.rs:38-83); every call is cut along the slices, whatever k, prefix or
    offset; points at infinity and the 72-byte stride travel with their slice.  Same bytes as the oracle."""
    ids = H.multi_device_ids(8)
    n = (1 << 15) + 321
    bases = orc.gen_bases(5, n)
    with cozk.Context(devices=ids) as mctx:
        srs = mctx.srs_register(bases, sliced=True)
        assert mctx.srs_len(srs) == n
        vecs = [orc.gen_scalars(d, 60 + i, n, stride=64) for i, d in enumerate(("uniform", "const", "wminus", "zero_half", "dup"))]
        want = [orc.msm(bases, v) for v in vecs]
        got = mctx.msm_batch(srs, vecs, n=n, stride=64)            # k > devices: still by point range
        for j in range(len(vecs)):
            assert (got[j] == want[j]).all(), j
        assert (mctx.msm_batch(srs, vecs[:1], n=n, stride=64)[0] == want[0]).all()
        for off, m in ((0, 1000), (17, 20000), (n - 5, 5), (n // len(ids) - 3, 7), (1, n - 1)):
            x = mctx.msm_batch(srs, [vecs[0]], n=m, base_offset=off, stride=64)
            assert (x[0] == orc.msm(bases[off:off + m], vecs[0][:m])).all(), (off, m)
        can = orc.gen_scalars("uniform", 8, n, form=1)
        assert (mctx.msm_batch(srs, can, form=1)[0] == orc.msm(bases, can, form=1)).all()
        with pytest.raises(cozk.CozkError) as e:
            mctx.msm_batch(srs, np.zeros((n + 1, 32), np.uint8))
        assert e.value.code == cozk.ERR_KEY_LENGTH
        d = mctx.testgen_scalars("uniform", 3, 64)
        with pytest.raises(cozk.CozkError):
            mctx.msm_batch_ptrs(srs, [d.ptr], 64, device=0)        # device-resident scalars: not with a sliced SRS
        d.free()
        mctx.srs_release(srs)
        b72 = np.zeros((n, 72), np.uint8)
        b72[:, :64] = bases
        inf = np.zeros(n, np.uint8)
        inf[3::7] = 1
        srs = mctx.srs_register(b72, infinity=inf, sliced=True)
        masked = vecs[0].copy()
        masked[3::7] = 0
        assert (mctx.msm_batch(srs, [vecs[0]], n=n, stride=64)[0] == orc.msm(bases, masked)).all()
        mctx.srs_release(srs)
        tiny = mctx.srs_register(bases[:2], sliced=True)            # fewer points than devices: empty slices are dropped
        assert (mctx.msm_batch(tiny, [vecs[0]], n=2, stride=64)[0] == orc.msm(bases[:2], vecs[0][:2])).all()
        mctx.srs_release(tiny)


def test_multi_pass_calls(cozk, orc):
    """A call longer than "max_points_per_pass" (2^26 by default) runs as several passes over point ranges whose partial
    results are added on the host: exercised here with a small pass length - batches, prefix / offset, host and device
    scalars, table and plain SRS."""
    n = 20_000
    bases = orc.gen_bases(19, n)
    with cozk.Context() as c2:
        c2.set_option("max_points_per_pass", 3000)
        for table in (True, False):
            c2.set_option("table_max_mib", 65536 if table else 0)
            srs = c2.srs_register(bases)
            vecs = [orc.gen_scalars(d, 500 + j, n) for j, d in enumerate(("uniform", "const", "wminus"))]
            got = c2.msm_batch(srs, vecs)
            for j, v in enumerate(vecs):
                assert (got[j] == orc.msm(bases, v)).all(), (table, j)
            x = c2.msm_batch(srs, [vecs[0]], n=7001, base_offset=999)
            assert (x[0] == orc.msm(bases[999:8000], vecs[0][:7001])).all()
            d = c2.alloc(n * 32).upload(vecs[0])
            assert (c2.msm_batch_ptrs(srs, [d.ptr], n, device=0)[0] == orc.msm(bases, vecs[0])).all()
            d.free()
            c2.srs_release(srs)


def test_ragged_batch(cozk, orc):
    """cozk_msm_ragged_device: vectors of very different lengths and offsets against one SRS in one launch (the shape of a
    PST13 opening's levels), with and without the SRS table, with points at infinity, both scalar forms; bit-exact with one
    oracle MSM per vector."""
    n_srs = 40_000
    bases = orc.gen_bases(17, n_srs)
    shapes = [(0, 20_000), (20_000, 10_000), (30_000, 5_000), (35_000, 2_500), (37_500, 1), (123, 31), (39_999, 1), (0, n_srs)]
    dists = ["uniform", "const", "wminus", "dup", "zero_half", "small16", "uniform", "uniform"]
    with cozk.Context() as c2:
        inf = np.zeros(n_srs, np.uint8)
        inf[7::11] = 1
        for table, flags in ((True, None), (False, None), (True, inf)):
            c2.set_option("table_max_mib", 65536 if table else 0)
            srs = c2.srs_register(bases, infinity=flags)
            for form in (0, 1):
                host = [orc.gen_scalars(d, 300 + j, ln, form=form) for j, (d, (_, ln)) in enumerate(zip(dists, shapes))]
                dev = [c2.alloc(h.nbytes).upload(h) for h in host]
                got = c2.msm_ragged(srs, [d.ptr for d in dev], [o for o, _ in shapes], [ln for _, ln in shapes], form=form)
                for j, ((off, ln), h) in enumerate(zip(shapes, host)):
                    sc = h.copy()
                    if flags is not None:
                        sc[flags[off:off + ln] != 0] = 0
                    assert (got[j] == orc.msm(bases[off:off + ln], sc, form=form)).all(), (table, flags is not None, form, j)
                for d in dev:
                    d.free()
            d = c2.alloc(64).upload(np.zeros(64, np.uint8))
            with pytest.raises(cozk.CozkError) as e:
                c2.msm_ragged(srs, [d.ptr], [n_srs - 1], [2])
            assert e.value.code == cozk.ERR_KEY_LENGTH
            d.free()
            c2.srs_release(srs)


def test_release_while_in_flight(cozk, ctx, orc):
    """cozk_srs_release on one thread while another is inside an MSM over the same SRS: the handle goes at
    once, the device memory only when the call in flight returns - its result is still exact."""
    import threading
    n = 1 << 16
    bases = orc.gen_bases(9, n)
    sc = orc.gen_scalars("uniform", 9, n)
    want = orc.msm(bases, sc)
    for _ in range(3):
        srs = ctx.srs_register(bases)
        res = {}

        def call():
            try:
                res["out"] = ctx.msm_batch(srs, sc)[0]
            except cozk.CozkError as e:  # released before the call looked the handle up
                res["err"] = e.code

        t = threading.Thread(target=call)
        t.start()
        ctx.srs_release(srs)
        t.join(timeout=120)
        assert ("out" in res and (res["out"] == want).all()) or res.get("err") == cozk.ERR_BAD_HANDLE, res
        with pytest.raises(cozk.CozkError):
            ctx.msm_batch(srs, sc)


@pytest.mark.parametrize("log2n,dists", [(22, ("uniform", "const", "wminus")), (24, ("uniform",))])
def test_oracle_parity_at_baseline_sizes(ctx, orc, log2n, dists):
    """Direct comparison with the CPU oracle at the sizes BASELINE.json names: 2^22 canonical scalars (configs[2], the
    form msm_bigint gets at co-noir-spartan/co-spartan/src/worker.rs:585/:804) in all three party distributions, and
    2^24 (configs[3]) - through the host-scalar call (streamed in chunks from 2^23 points) and the device-resident one."""
    n = 1 << log2n
    dbases = ctx.testgen_bases(1, n)
    bases = dbases.download().reshape(n, 64)
    # the device generator is the oracle's generator: spot-check both ends
    assert (bases[:2048] == orc.gen_bases(1, 2048)).all()
    assert (bases[n - 1024:] == orc.gen_bases(1, 1024, start=n - 1024)).all()
    srs = ctx.srs_register_device(dbases, n)
    dbases.free()
    for dist in dists:
        sc = orc.gen_scalars(dist, 2, n, form=1)
        want = orc.msm(bases, sc, form=1)
        got = ctx.msm_batch(srs, sc, form=1)[0]
        assert (got == want).all(), (log2n, dist, "host scalars")
        d = ctx.alloc(n * 32).upload(sc)
        got = ctx.msm_batch_ptrs(srs, [d.ptr], n, form=1, device=0)[0]
        d.free()
        assert (got == want).all(), (log2n, dist, "device scalars")
    ctx.srs_release(srs)


def test_precomputed_table_matches_plain_path(cozk, orc):
    """An SRS registered with a 2^(c*w) * P table (default) and one without give the same bytes; prefix / offset calls
    and batches included; the stats show which path ran."""
    n = 1 << 14
    bases = orc.gen_bases(6, n)
    with cozk.Context() as c2:
        with_table = c2.srs_register(bases)
        c2.set_option("table_max_mib", 0)
        plain = c2.srs_register(bases)
        c2.set_option("table_max_mib", 65536)
        vecs = [orc.gen_scalars(d, 70 + i, n, stride=64) for i, d in enumerate(("uniform", "const", "wminus", "dup"))]
        a = c2.msm_batch(with_table, vecs, n=n, stride=64)
        wa = c2.last_stats()["windows"]
        b = c2.msm_batch(plain, vecs, n=n, stride=64)
        wb = c2.last_stats()["windows"]
        assert (a == b).all()
        for j, v in enumerate(vecs):
            assert (a[j] == orc.msm(bases, v)).all(), j
        assert c2.last_stats()["window"] >= 2 and wa <= wb
        # prefix and mid-SRS slice
        for (off, m) in ((0, n // 2), (1000, 9000), (n - 40, 40)):
            x = c2.msm_batch(with_table, [vecs[0]], n=m, base_offset=off, stride=64)
            assert (x[0] == orc.msm(bases[off:off + m], vecs[0][:m])).all(), (off, m)
        small = orc.gen_scalars("small16", 5, n)
        assert (c2.msm_batch(with_table, small, max_num_bits=16)[0] == orc.msm(bases, small)).all()
        c2.srs_release(with_table)
        c2.srs_release(plain)


def test_dominant_digit_path(cozk, orc):
    """Whole-SRS calls on vectors whose windows are dominated by one digit (the real co-jolt share shapes: constant
    vectors, w - c0 - c1, zero padding) use the precomputed row totals: same bytes as the plain layout and the oracle, far
    fewer (key, val) pairs; uniform vectors keep the plain layout; prefix calls and SRSs with points at infinity too."""
    n = 1 << 15
    bases = orc.gen_bases(12, n)
    dists = ("const", "wminus", "zero_half", "uniform", "dup", "small16")
    vecs = [orc.gen_scalars(d, 110 + i, n, stride=64) for i, d in enumerate(dists)]
    broken = vecs[0].copy()
    broken[5] = vecs[3][5]
    broken[n - 3] = 0
    vecs.append(broken)
    want = np.stack([orc.msm(bases, v) for v in vecs])
    with cozk.Context() as c2:
        c2.set_option("dominant_min_points", 0)
        with_table = c2.srs_register(bases)
        c2.set_option("table_max_mib", 0)
        plain = c2.srs_register(bases)
        for srs in (with_table, plain):
            pairs = {}
            for dom in (1, 0):
                c2.set_option("dominant", dom)
                for j, v in enumerate(vecs):
                    assert (c2.msm_batch(srs, [v], stride=64)[0] == want[j]).all(), (srs, dom, j)
                    pairs[(dom, j)] = c2.last_stats()["pairs"]
                assert (c2.msm_batch(srs, vecs, stride=64) == want).all(), (srs, dom, "batch")
            windows = c2.last_stats()["windows"]
            assert pairs[(1, 0)] <= windows + 1 and pairs[(0, 0)] == windows * n      # constant vector: the row totals only
            assert pairs[(1, 1)] < 0.6 * pairs[(0, 1)]                                  # w - c0 - c1: the high windows collapse
            assert pairs[(1, 2)] < 0.6 * pairs[(0, 2)]                                  # half zeros: found through the tail sample
            assert pairs[(1, 3)] == pairs[(0, 3)]                                       # uniform: plain layout
            assert pairs[(1, 6)] <= windows * 5 + 1                                     # two scalars off the pattern: 2 pairs each
            c2.set_option("dominant", 1)
            # exact halvings of the SRS carry their own totals (a 2^16 polynomial against a 2^22 SRS): dominant path too;
            # any other prefix and any offset keep the plain layout; same bytes everywhere
            for m_, off in ((n // 2, 0), (n // 8, 0), (n // 2 + 1, 0), (n // 4, n // 4)):
                got = c2.msm_batch(srs, [vecs[0]], n=m_, base_offset=off, stride=64)[0]
                assert (got == orc.msm(bases[off:off + m_], vecs[0][:m_])).all(), (srs, m_, off)
                made = c2.last_stats()["pairs"]
                assert (made <= c2.last_stats()["windows"] + 1) == (off == 0 and m_ in (n // 2, n // 8)), (m_, off, made)
        # canonical form, dense stride
        can = orc.gen_scalars("const", 130, n, form=1)
        assert (c2.msm_batch(with_table, can, form=1)[0] == orc.msm(bases, can, form=1)).all()
        inf = np.zeros(n, np.uint8)
        inf[::5] = 1
        srs_inf = c2.srs_register(bases, infinity=inf)
        masked = vecs[0].copy()
        masked[::5] = 0
        assert (c2.msm_batch(srs_inf, [vecs[0]], stride=64)[0] == orc.msm(bases, masked)).all()
        for h in (with_table, plain, srs_inf):
            c2.srs_release(h)


def test_randomised_shapes(ctx, orc):
    """Seeded random shapes through the C ABI: n, base_offset, batch size, stride, form, distribution, bit hint."""
    rng = np.random.default_rng(20261018)
    n_srs = 6000
    bases = orc.gen_bases(11, n_srs)
    srs = ctx.srs_register(bases)
    dists = list(pyref.DISTS)
    for trial in range(24):
        n = int(rng.integers(1, 3000))
        off = int(rng.integers(0, n_srs - n + 1))
        k = int(rng.integers(1, 6))
        stride = int(rng.choice([32, 64, 48]))
        form = int(rng.integers(0, 2))
        vecs, want = [], []
        for j in range(k):
            d = dists[int(rng.integers(0, len(dists)))]
            v = orc.gen_scalars(d, 1000 + 10 * trial + j, n, form=form, stride=stride)
            vecs.append(v)
            want.append(orc.msm(bases[off:off + n], v, form=form))
        got = ctx.msm_batch(srs, vecs, n=n, base_offset=off, stride=stride, form=form)
        for j in range(k):
            assert (got[j] == want[j]).all(), (trial, n, off, k, stride, form, j)
    ctx.srs_release(srs)


def test_concurrent_callers(cozk, ctx, orc):
    """Two host threads calling into one context at once (the reference runs two commits under rayon::join,
    co-jolt/src/jolt/vm/jolt/witness.rs:350-365): calls serialise per device and both results are exact."""
    import threading
    n = 1 << 14
    bases = orc.gen_bases(1, n)
    srs = ctx.srs_register(bases)
    vecs = [orc.gen_scalars(d, 90 + i, n) for i, d in enumerate(("uniform", "const", "wminus", "dup"))]
    want = [orc.msm(bases, v) for v in vecs]
    results, errors = {}, []

    def worker(idx):
        try:
            for rep in range(6):
                j = (idx + rep) % len(vecs)
                out = ctx.msm_batch(srs, vecs[j])
                if not (out[0] == want[j]).all():
                    errors.append((idx, rep, j))
            results[idx] = True
        except Exception as e:  # noqa: BLE001
            errors.append((idx, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert len(results) == 3
    ctx.srs_release(srs)


def test_large_batch_of_tiny_vectors(ctx, orc):
    """1000 vectors of 8 points: tens of thousands of (vector, window) pairs in one group (the last PST13 opening levels
    and small lookup tables produce such shapes)."""
    n, k = 8, 1000
    bases = orc.gen_bases(13, n)
    srs = ctx.srs_register(bases)
    vecs = [orc.gen_scalars("uniform" if j % 3 else "const", 2000 + j, n) for j in range(k)]
    got = ctx.msm_batch(srs, vecs, n=n)
    for j in range(0, k, 37):
        assert (got[j] == orc.msm(bases, vecs[j])).all(), j
    assert (got[k - 1] == orc.msm(bases, vecs[k - 1])).all()
    ctx.srs_release(srs)


def test_streamed_host_vector(cozk, orc):
    """A long host-resident vector is streamed in point chunks through one bucket set (H2D of chunk i+1 overlaps the
    compute of chunk i): same bytes as the one-shot path, with and without the SRS table, Rep3 stride, odd lengths."""
    n = (1 << 16) + 777
    bases = orc.gen_bases(21, n)
    with cozk.Context() as c2:
        c2.set_option("stream_min_points", 1 << 12)
        srs_t = c2.srs_register(bases)
        c2.set_option("table_max_mib", 0)
        srs_p = c2.srs_register(bases)
        for dist in ("uniform", "const", "wminus"):
            sc = orc.gen_scalars(dist, 31, n, stride=64)
            want = orc.msm(bases, sc)
            for chunks in (0, 2, 4, 7):  # 0 = chosen from the vector length
                c2.set_option("stream_chunks", chunks)
                assert (c2.msm_batch(srs_t, [sc], n=n, stride=64)[0] == want).all(), (dist, chunks, "table")
                assert (c2.msm_batch(srs_p, [sc], n=n, stride=64)[0] == want).all(), (dist, chunks, "plain")
        sc = orc.gen_scalars("uniform", 32, n)
        c2.set_option("stream_chunks", 3)
        got = c2.msm_batch(srs_t, [sc], n=50001, base_offset=1234)
        assert (got[0] == orc.msm(bases[1234:1234 + 50001], sc[:50001])).all()
        c2.srs_release(srs_t)
        c2.srs_release(srs_p)
        # a vector the dominant-digit mode applies to (whole SRS, constant share) is recognised from its head and takes the
        # one-shot path with the row totals instead of being streamed; a uniform one is streamed
        m2 = 1 << 15
        c2.set_option("table_max_mib", 65536)
        c2.set_option("dominant_min_points", 0)
        c2.set_option("stream_chunks", 2)
        srs2 = c2.srs_register(bases[:m2])
        for dist, few in (("const", True), ("wminus", False), ("uniform", False)):
            v = orc.gen_scalars(dist, 41, m2)
            assert (c2.msm_batch(srs2, [v])[0] == orc.msm(bases[:m2], v)).all(), dist
            st = c2.last_stats()
            assert (st["pairs"] <= st["windows"] + 1) == few, (dist, st["pairs"])
            if dist == "wminus":
                assert st["pairs"] < 0.6 * st["windows"] * m2
        c2.srs_release(srs2)


def test_row_column_bucket_reduce(cozk, orc):
    """Option reduce_2d: the bucket reduce as row sums + column sums + one masked sum per index bit (two tree sums deep)
    instead of group running sums.  Every window size, table and per-window buckets, batches (on-device finish), a chunked
    host call, a resident device vector: the same bytes as the oracle's."""
    n = 3000
    bases = orc.gen_bases(2, n)
    with cozk.Context() as c2:
        c2.set_option("reduce_2d", 1)
        srs = c2.srs_register(bases)
        u, c0 = orc.gen_scalars("uniform", 4, n), orc.gen_scalars("const", 4, n)
        wu, wc = orc.msm(bases, u), orc.msm(bases, c0)
        assert (c2.msm_batch(srs, u)[0] == wu).all()  # table mode
        for c in range(2, 23):
            c2.set_option("window", c)
            assert (c2.msm_batch(srs, u)[0] == wu).all(), c
            assert (c2.msm_batch(srs, c0)[0] == wc).all(), c
        c2.set_option("window", 0)
        vecs = [orc.gen_scalars(d, 70 + j, n) for j, d in enumerate(("uniform", "const", "wminus", "dup", "zero_half", "small16", "uniform"))]
        got = c2.msm_batch(srs, vecs, n=n)  # 7 vectors: finish on the device
        for j, v in enumerate(vecs):
            assert (got[j] == orc.msm(bases, v)).all(), j
        c2.srs_release(srs)
        m = 1 << 17
        bases2 = orc.gen_bases(5, m)
        srs2 = c2.srs_register(bases2)
        c2.set_option("stream_min_points", 1 << 14)
        for dist in ("uniform", "wminus"):
            v = orc.gen_scalars(dist, 9, m)
            want = orc.msm(bases2, v)
            assert (c2.msm_batch(srs2, [v])[0] == want).all(), dist  # chunked host call
            dv = c2.alloc(v.nbytes)
            dv.upload(v)
            assert (c2.msm_batch_ptrs(srs2, [dv.ptr], m, device=0)[0] == want).all(), dist
            dv.free()
        c2.srs_release(srs2)
