"""CPU tier: the kernel thread bodies (co-zkvms_b200/csrc/msm_kernels.cuh), compiled for the host and run as
sequential loops, against the oracle.  This checks the pipeline LOGIC without a GPU; the PTX field arithmetic is
checked by test_field_ptx.py and, on the device, by the gpu tier."""
import numpy as np
import pytest

from oracle import pyref
from tests import emul
from tests import helpers as H


def test_field_bodies(orc):
    rng = np.random.default_rng(1)
    vals = [0, 1, H.P - 1, H.P - 2, (1 << 253)] + [int.from_bytes(rng.bytes(32), "little") % H.P for _ in range(40)]
    a = np.stack([H.le32(v) for v in vals])
    b = np.roll(a, 3, axis=0)
    for op in ("mul", "add", "sub"):
        want = orc.field_op("fq", op, a.view(np.uint64), b.view(np.uint64)).view(np.uint8)
        assert (emul.field_op(op, a, b) == want).all(), op
    assert (emul.field_op("sqr", a) == orc.field_op("fq", "sqr", a.view(np.uint64)).view(np.uint8)).all()
    assert (emul.field_op("inv", a) == orc.field_op("fq", "inv", a.view(np.uint64)).view(np.uint8)).all()
    ar = np.stack([H.le32(v % H.R) for v in vals])
    assert (emul.field_op("fr_from_mont", ar) == orc.field_op("fr", "from_mont", ar.view(np.uint64)).view(np.uint8)).all()


def test_group_law_bodies(orc):
    g = H.golden()
    for case in g["adds"]:
        a, b = H.point_wire(H.parse_point(case["a"])), H.point_wire(H.parse_point(case["b"]))
        want = H.point_wire(H.parse_point(case["sum"]))
        assert (emul.g1_op("add", a, b)[0] == want).all(), case
        if case["b"] is not None:
            assert (emul.g1_op("madd", a, b)[0] == want).all(), case
    pts = np.zeros((16, 72), np.uint8)
    pts[:, :64] = orc.gen_bases(2, 16)
    assert (emul.g1_op("dbl", pts) == orc.g1_op("dbl", pts)).all()
    assert (emul.g1_op("add", pts, np.roll(pts, 1, axis=0)) == orc.g1_op("add", pts, np.roll(pts, 1, axis=0))).all()


def test_window_choice_respects_bucket_budget():
    """Large batches: every window that is considered keeps g * W * 2^(c-1) buckets inside the budget whatever the stack
    held before the call (cost[] entries behind the first misfit were read unwritten)."""
    budget = 1 << 25
    for g in (1, 3, 54, 138, 256, 1000, 4096):
        for n in (1 << 10, 1 << 16, 1 << 20):
            seen = set()
            for poison in (0.0, -1.0, 1e300, 5.0):
                c, gfit = emul.choose_window(n, g, 254, budget, poison)
                seen.add((c, gfit))
                if c:
                    W = (254 + 1 + c - 1) // c
                    assert g * W * (1 << (c - 1)) <= budget, (g, n, c)
                assert 1 <= gfit <= g
                cf, _ = emul.choose_window(n, gfit, 254, budget, poison)
                Wf = (254 + 1 + cf - 1) // cf
                assert cf and gfit * Wf * (1 << (cf - 1)) <= budget
            assert len(seen) == 1, (g, n, seen)
    assert emul.choose_window(1 << 16, 1 << 26, 254, budget)[0] == 0


def test_msm_golden():
    for case in H.golden()["msm"]:
        if case["n"] > 300:
            continue
        pts, sc = H.golden_msm_inputs(case)
        want = H.point_wire(H.parse_point(case["result"]))
        for form in (0, 1):
            got, _ = emul.msm(H.bases_wire(pts), H.scalars_wire(sc, form), form=form)
            assert (got[0] == want).all(), (case["n"], case["dist"], form)


@pytest.mark.parametrize("dist", pyref.DISTS)
def test_msm_all_windows(orc, dist):
    """Every window size exercises a different level structure of the accumulate / reduce trees."""
    n = 200
    bases = orc.gen_bases(1, n)
    sc = orc.gen_scalars(dist, 5, n)
    want = orc.msm(bases, sc)
    for c in (2, 3, 4, 6, 9, 12, 14):
        got, st = emul.msm(bases, sc, c=c)
        assert (got[0] == want).all(), (dist, c, st)


def test_msm_batch_strides_bits_infinity(orc):
    n, g = 150, 3
    bases = orc.gen_bases(7, n)
    vecs = [orc.gen_scalars(d, 20 + i, n, stride=64) for i, d in enumerate(("uniform", "const", "wminus"))]
    got, _ = emul.msm(bases, np.concatenate(vecs), g=g, stride=64)
    for j in range(g):
        assert (got[j] == orc.msm(bases, vecs[j])).all(), j
    # max_num_bits hint drops high windows; result unchanged
    small = orc.gen_scalars("small16", 3, n)
    got, st = emul.msm(bases, small, bits=16)
    assert st[1] <= 6 and (got[0] == orc.msm(bases, small)).all()
    # points at infinity contribute nothing
    inf = np.zeros(n, np.uint8)
    inf[::3] = 1
    u = orc.gen_scalars("uniform", 4, n)
    masked = u.copy()
    masked[::3] = 0
    got, _ = emul.msm(bases, u, infinity=inf)
    assert (got[0] == orc.msm(bases, masked)).all()


def test_accumulate_chunk_lengths(orc):
    """The level-1 chunk length is chosen per call so that the last wave of threads is full: every length gives the same
    point, for runs much longer and much shorter than a chunk."""
    n = 3000
    bases = orc.gen_bases(4, n)
    vecs = [orc.gen_scalars(d, 40 + i, n) for i, d in enumerate(("uniform", "const", "wminus", "zero_half"))]
    want = [orc.msm(bases, v) for v in vecs]
    try:
        for force_l in (4, 16, 17, 29, 35, 40, 256):
            emul.set_acc_chunk(0, force_l)
            for v, w in zip(vecs, want):
                for kw in ({"c": 5}, {"table_c": 9}):
                    got, _ = emul.msm(bases, v, **kw)
                    assert (got[0] == w).all(), (force_l, kw)
        for resident in (64, 1000, 75776):
            emul.set_acc_chunk(resident, 0)
            got, _ = emul.msm(bases, vecs[0], c=7)
            assert (got[0] == want[0]).all(), resident
    finally:
        emul.set_acc_chunk(0, 0)


@pytest.mark.parametrize("dist", ["const", "wminus", "zero_half", "uniform", "dup", "small16"])
def test_dominant_digit_path(orc, dist):
    """Windows in which (nearly) every scalar has the same digit - the real co-jolt share shapes - run as
    sum_{d_i != c} d_i T_i + c (S - sum_{d_i != c} T_i) with the row total S precomputed; zero-dominated windows drop
    their zero digits.  Same point as the plain layout and the oracle, with the table and with per-window buckets, for a
    batch, and for vectors in which a few scalars break the pattern."""
    n = 600
    bases = orc.gen_bases(5, n)
    sc = orc.gen_scalars(dist, 50, n)
    broken = sc.copy()
    broken[7] = orc.gen_scalars("uniform", 51, 1)[0]      # one scalar off the pattern
    broken[n - 1] = 0
    vecs = [sc, broken, orc.gen_scalars("const", 52, n)]
    want = [orc.msm(bases, v) for v in vecs]
    try:
        emul.set_dominant(True)
        for kw in ({"c": 6}, {"table_c": 7}, {"table_c": 11}, {"c": 13}):
            for v, w in zip(vecs, want):
                got, st = emul.msm(bases, v, **kw)
                assert (got[0] == w).all(), (dist, kw)
            assert st[5] == 1 and st[4] <= st[1] + 1      # the constant vector: one pair (the row total) per window
            got, _ = emul.msm(bases, np.concatenate(vecs), g=3, **kw)
            assert all((got[j] == want[j]).all() for j in range(3)), (dist, kw, "batch")
        # element 0 itself off the pattern: the candidate is not the dominant digit, the result is still right
        odd = sc.copy()
        odd[0] = orc.gen_scalars("uniform", 53, 1)[0]
        got, _ = emul.msm(bases, odd, table_c=9)
        assert (got[0] == orc.msm(bases, odd)).all()
        # a prefix of the SRS (a short polynomial against a long SRS), with the table and with per-window buckets
        for kw in ({"table_c": 8}, {"c": 7}):
            got, st = emul.msm(bases, vecs[2][: n // 2], n=n // 2, **kw)
            assert st[5] == 1 and (got[0] == orc.msm(bases[: n // 2], vecs[2][: n // 2])).all(), kw
        got, st = emul.msm(bases, orc.gen_scalars("uniform", 54, n), table_c=9)
        assert st[5] == 0 and st[4] == st[1] * n          # uniform scalars: nothing dominant, the plain layout
    finally:
        emul.set_dominant(False)


def test_degenerate_many_levels(orc):
    """All points in one bucket per window (co-jolt party 0/1 shares): the open-run / partial-merge path, 3 levels deep."""
    n = 20000
    bases = orc.gen_bases(1, n)
    sc = orc.gen_scalars("const", 8, n)
    got, st = emul.msm(bases, sc, c=8)
    assert st[2] >= 4
    assert (got[0] == orc.msm(bases, sc)).all()
    n = 1100
    bases, sc = bases[:n], sc[:n]
    # duplicated bases with equal scalars force P + P inside a bucket
    dup = bases.copy()
    dup[1::2] = dup[0::2]
    got, _ = emul.msm(dup, sc, c=5)
    assert (got[0] == orc.msm(dup, sc)).all()
    # P + (-P): bucket collapses to the identity
    neg = np.tile(H.le32(H.R - 1), (n, 1))
    neg[0::2] = H.le32(1)
    got, _ = emul.msm(dup, neg, form=1, c=4)
    assert got[0][64] == 1


def test_radix29_experiment_is_correct(orc):
    """experiments/radix29 (not in the product): 9 x 29-bit field and its lazy mixed addition, against the oracle."""
    rng = np.random.default_rng(5)
    vals = [0, 1, 2, H.P - 1, H.P - 2, 1 << 253, H.P >> 1, (1 << 29) - 1, 1 << 232]
    vals += [int.from_bytes(rng.bytes(32), "little") % H.P for _ in range(200)]
    a = np.stack([H.le32(v) for v in vals])
    b = np.roll(a, 5, axis=0)
    for op, o in (("f29_mul", "mul"), ("f29_add", "add"), ("f29_sub", "sub")):
        assert (emul.field_op(op, a, b) == orc.field_op("fq", o, a.view(np.uint64), b.view(np.uint64)).view(np.uint8)).all(), op
    assert (emul.field_op("f29_sqr", a) == orc.field_op("fq", "sqr", a.view(np.uint64)).view(np.uint8)).all()
    assert (emul.field_op("f29_inv", a[:30]) == orc.field_op("fq", "inv", a[:30].view(np.uint64)).view(np.uint8)).all()
    n = 120
    pts = np.zeros((n, 72), np.uint8)
    pts[:, :64] = orc.gen_bases(2, n)
    other = np.roll(pts, 1, axis=0)
    other[::7] = pts[::7]
    other[3::11] = orc.g1_op("neg", pts[3::11])
    pts[5::13] = H.point_wire(None)
    assert (emul.g1_op("f29_madd3", pts, other) == orc.g1_op("add", pts, other)).all()


def test_precomputed_table_mode(orc):
    """SRS table of 2^(c*w) * P rows: all windows share one bucket set; prefix and offset calls index into the rows."""
    n_srs = 300
    bases = orc.gen_bases(4, n_srs)
    for dist in ("uniform", "const", "wminus", "dup"):
        sc = orc.gen_scalars(dist, 6, n_srs)
        for tc in (3, 7, 12):
            got, st = emul.msm(bases, sc, table_c=tc)
            assert (got[0] == orc.msm(bases, sc)).all(), (dist, tc)
    sc = orc.gen_scalars("uniform", 7, 120)
    got, _ = emul.msm(bases, sc, table_c=6, n=120, base_offset=50)
    assert (got[0] == orc.msm(bases[50:170], sc)).all()
    # a batch, Rep3 stride, and a small-scalar hint (fewer table rows used)
    vecs = [orc.gen_scalars(d, 30 + i, n_srs, stride=64) for i, d in enumerate(("uniform", "const"))]
    got, _ = emul.msm(bases, np.concatenate(vecs), g=2, stride=64, table_c=9)
    for j in range(2):
        assert (got[j] == orc.msm(bases, vecs[j])).all()
    small = orc.gen_scalars("small16", 3, n_srs)
    got, st = emul.msm(bases, small, bits=16, table_c=8)
    assert st[1] == 3 and (got[0] == orc.msm(bases, small)).all()


def test_streamed_chunks_merge_into_one_bucket_set(orc):
    """One vector fed in point chunks (the engine's H2D-overlap mode): later chunks ADD to the buckets of earlier ones."""
    n = 700
    bases = orc.gen_bases(9, n)
    for dist in ("uniform", "const", "wminus", "zero_half"):
        sc = orc.gen_scalars(dist, 21, n)
        want = orc.msm(bases, sc)
        for chunks, c, tc in ((2, 0, 0), (4, 6, 0), (5, 0, 7), (3, 0, 11)):
            got, _ = emul.msm(bases, sc, c=c, table_c=tc, stream_chunks=chunks)
            assert (got[0] == want).all(), (dist, chunks, c, tc)
    inf = np.zeros(n, np.uint8)
    inf[::5] = 1
    u = orc.gen_scalars("uniform", 4, n)
    masked = u.copy()
    masked[::5] = 0
    got, _ = emul.msm(bases, u, infinity=inf, stream_chunks=3)
    assert (got[0] == orc.msm(bases, masked)).all()


@pytest.mark.parametrize("dist", ["uniform", "const", "dup", "zero_half"])
def test_row_column_bucket_reduce(orc, dist):
    """The row / column form of the bucket reduce (two tree sums instead of group running sums + masked sums) for every
    window size - odd and even bucket-index widths, one bucket - with and without a table, single vector and batch."""
    n = 260
    bases = orc.gen_bases(3, n)
    sc = orc.gen_scalars(dist, 8, n)
    want = orc.msm(bases, sc)
    try:
        emul.set_reduce_2d(1)
        for c in (2, 3, 4, 5, 8, 11, 14):
            got, st = emul.msm(bases, sc, c=c)
            assert (got[0] == want).all(), (dist, c)
        for tc in (3, 6, 9):
            got, _ = emul.msm(bases, sc, table_c=tc)
            assert (got[0] == want).all(), (dist, tc)
        vecs = [orc.gen_scalars(d, 50 + i, n, stride=64) for i, d in enumerate(("uniform", "wminus", "small16"))]
        got, _ = emul.msm(bases, np.concatenate(vecs), g=3, stride=64, table_c=7)
        for j in range(3):
            assert (got[j] == orc.msm(bases, vecs[j])).all(), j
    finally:
        emul.set_reduce_2d(0)


def test_table_build_in_two_steps(orc):
    """SRS table: the engine's chain + batch-normalisation kernels (one inversion per point) give the bytes of the row-by-row
    contract body (one inversion per row and point), for slabs that do and do not divide the SRS."""
    n = 37
    bases = orc.gen_bases(5, n)
    for c, rows in ((3, 5), (7, 4), (13, 20)):
        want = emul.build_table(bases, c, rows)
        for slab in (n, 16, 5):
            assert (emul.build_table(bases, c, rows, mode=1, slab=slab) == want).all(), (c, rows, slab)
    # row w really is 2^(c w) P
    pts = [(pyref.from_mont(H.to_int(r[:32]), H.P), pyref.from_mont(H.to_int(r[32:]), H.P)) for r in bases[:3]]
    t = emul.build_table(bases, 7, 4, mode=1)
    for i, p in enumerate(pts):
        q = pyref.mul(1 << 14, p)
        row = t[2 * n + i]
        assert (pyref.from_mont(H.to_int(row[:32]), H.P), pyref.from_mont(H.to_int(row[32:]), H.P)) == q
