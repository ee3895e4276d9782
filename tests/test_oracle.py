"""CPU tier: the oracle (oracle/bn254.c) against the golden fixtures, public known answers and identities."""
import numpy as np
import pytest

from oracle import pyref
from tests import helpers as H


def test_known_answer_2G(orc):
    # EIP-196 / every BN254 library: 2*(1,2)
    two = orc.wire_to_point(orc.g1_op("dbl", H.point_wire(pyref.G1)[None, :])[0])
    assert two == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                   9918110051302171585080402603319702774565515993150576347155970296011118125764)


def test_generator_multiples(orc):
    G = H.point_wire(pyref.G1)
    for kat in H.golden()["generator_multiples"]:
        k = int(kat["k"], 16)
        got = orc.g1_mul(G, H.le32(k % (1 << 256)), form=orc.CANON) if k < (1 << 256) else None
        assert (got == H.point_wire(H.parse_point(kat["kG"]))).all(), kat["k"]


def test_field_golden(orc):
    for f in H.golden()["field"]:
        mod = H.P if f["field"] == "fq" else H.R
        mont = (lambda x: H.le32(x * pyref.MONT_R % mod))
        a, b = int(f["a"], 16), int(f["b"], 16)
        A = np.frombuffer(mont(a).tobytes(), dtype=np.uint64)[None, :]
        B = np.frombuffer(mont(b).tobytes(), dtype=np.uint64)[None, :]
        for op, key in (("mul", "mul"), ("add", "add"), ("sub", "sub")):
            got = orc.field_op(f["field"], op, A, B)
            assert (got.view(np.uint8).reshape(-1) == mont(int(f[key], 16))).all(), (f["field"], op)
        got = orc.field_op(f["field"], "inv", A)
        assert (got.view(np.uint8).reshape(-1) == mont(int(f["inv_a"], 16))).all()
        # Montgomery round trip
        canon = np.frombuffer(H.le32(a).tobytes(), dtype=np.uint64)[None, :]
        assert (orc.field_op(f["field"], "to_mont", canon) == A).all()
        assert (orc.field_op(f["field"], "from_mont", A) == canon).all()


def test_generators_match_golden(orc):
    g = H.golden()
    b = orc.gen_bases(1, 8)
    assert (b == H.bases_wire([H.parse_point(p) for p in g["bases_seed1"]])).all()
    s = orc.gen_scalars("uniform", 2, 8, form=orc.CANON)
    assert [H.to_int(s[i]) for i in range(8)] == [int(x, 16) for x in g["scalars_seed2_uniform"]]
    for dist in pyref.DISTS:
        s = orc.gen_scalars(dist, 9, 40, form=orc.MONT)
        assert [pyref.from_mont(H.to_int(s[i]), H.R) for i in range(40)] == pyref.scalars(dist, 9, 40), dist


def test_adds_golden(orc):
    for case in H.golden()["adds"]:
        a, b = H.point_wire(H.parse_point(case["a"])), H.point_wire(H.parse_point(case["b"]))
        got = orc.g1_op("add", a[None, :], b[None, :])[0]
        assert (got == H.point_wire(H.parse_point(case["sum"]))).all(), case
        assert orc.g1_is_valid(got)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_msm_golden(orc, threads):
    for case in H.golden()["msm"]:
        pts, sc = H.golden_msm_inputs(case)
        want = H.point_wire(H.parse_point(case["result"]))
        for form in (0, 1):
            got = orc.msm(H.bases_wire(pts), H.scalars_wire(sc, form), form=form, threads=threads)
            assert (got == want).all(), (case["n"], case["dist"], form)
    # stride 64 (Rep3 AoS share a) reads the same values
    case = H.golden()["msm"][10]
    pts, sc = H.golden_msm_inputs(case)
    got = orc.msm(H.bases_wire(pts), H.scalars_wire(sc, 0, stride=64), threads=threads)
    assert (got == H.point_wire(H.parse_point(case["result"]))).all()


def test_pippenger_matches_naive(orc):
    for n in (1, 7, 31, 32, 100, 300):
        bases = orc.gen_bases(3, n)
        for dist in ("uniform", "const", "wminus", "dup", "zero_half"):
            sc = orc.gen_scalars(dist, 11, n)
            assert (orc.msm(bases, sc) == orc.msm_naive(bases, sc)).all(), (n, dist)


def test_identities(orc):
    n = 2000
    bases = orc.gen_bases(5, n)
    s = orc.gen_scalars("uniform", 1, n, form=orc.CANON)
    t = orc.gen_scalars("uniform", 2, n, form=orc.CANON)
    # linearity: MSM(s + t) = MSM(s) + MSM(t)   (the reference's own property test, pst13.rs:536)
    st = np.zeros_like(s)
    for i in range(n):
        st[i] = H.le32((H.to_int(s[i]) + H.to_int(t[i])) % H.R)
    lhs = orc.msm(bases, st, form=orc.CANON)
    rhs = orc.g1_op("add", orc.msm(bases, s, form=orc.CANON)[None, :], orc.msm(bases, t, form=orc.CANON)[None, :])[0]
    assert (lhs == rhs).all()
    # split = sum of parts (combine_comm, snarks-core/src/poly/commitment.rs:56-63)
    a = orc.msm(bases[:700], s[:700], form=orc.CANON)
    b = orc.msm(bases[700:], s[700:], form=orc.CANON)
    assert (orc.g1_op("add", a[None, :], b[None, :])[0] == orc.msm(bases, s, form=orc.CANON)).all()
    # constant scalar c: MSM = c * sum(P_i)   (exactly the co-jolt party 0/1 share shape)
    c = orc.gen_scalars("const", 4, n, form=orc.CANON)
    ones = np.zeros_like(c)
    ones[:, 0] = 1
    assert (orc.msm(bases, c, form=orc.CANON) == orc.g1_mul(orc.msm(bases, ones, form=orc.CANON), c[0], form=orc.CANON)).all()
    # r - 1 everywhere: -sum(P_i)
    m1 = np.tile(H.le32(H.R - 1), (n, 1))
    assert (orc.msm(bases, m1, form=orc.CANON) == orc.g1_op("neg", orc.msm(bases, ones, form=orc.CANON)[None, :])[0]).all()
    # empty input
    assert orc.msm(bases[:0], s[:0], n=0)[64] == 1


def test_window_rule(orc):
    # ark-ec: c = 3 below 32 points, else ceil(log2 n) * 69 / 100 + 2
    assert orc.msm_window(31) == 3
    assert orc.msm_window(32) == 5
    assert orc.msm_window(1 << 16) == 13
    assert orc.msm_window(1 << 20) == 15
    assert orc.msm_window(1 << 24) == 18
