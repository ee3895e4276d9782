"""CPU tier for SURVEY.md 8(f) rows N1 / N2 / N4: the thread bodies of co-zkvms_b200/csrc/rep3_kernels.cuh, compiled
for the host and run as loops, against the Python restatement (oracle/rep3ref.py)."""
import numpy as np
import pytest

from oracle import pyref, rep3ref
from tests import emul
from tests import helpers as H

R = H.R


def _rand_fr(seed, n):
    return [pyref.scalar_uniform(seed, i) for i in range(n)]


def _shared_mont(vals):
    out = np.zeros((len(vals), 64), np.uint8)
    for i, (a, b) in enumerate(vals):
        out[i, :32] = H.fr_mont(a)
        out[i, 32:] = H.fr_mont(b)
    return out


def _from_shared(arr):
    return [(pyref.from_mont(H.to_int(row[:32]), R), pyref.from_mont(H.to_int(row[32:]), R)) for row in arr]


def test_wire_roundtrip_python_side():
    coeffs = list(zip(_rand_fr(1, 8), _rand_fr(2, 8)))
    for tagged in (False, True):
        raw = rep3ref.serialize_rep3_dense(coeffs, tagged=tagged)
        assert len(raw) == (1 if tagged else 0) + 8 + 8 + 8 * 64 + 8 + 1 + 8 + 16
        got, used = rep3ref.deserialize_rep3_dense(raw, tagged=tagged)
        assert got == coeffs and used == len(raw)
    raw = rep3ref.serialize_rep3_dense(coeffs, bound_coeffs=coeffs[:2], scratch=coeffs[:3], chunk_range=(2, 6))
    got, used = rep3ref.deserialize_rep3_dense(raw + b"trailing")
    assert got == coeffs[2:6] and used == len(raw)
    with pytest.raises(ValueError):
        rep3ref.deserialize_rep3_dense(raw[:-1])


def test_lazy_accumulator_extremes():
    """576-bit accumulator and its single reduction: (r-1)^2 terms, enough of them to reach the two top limbs."""
    top = np.stack([H.le32(R - 1)] * 64)
    got = emul.wide_dot(top, top, repeat=300)
    want = (R - 1) * (R - 1) * 64 * 300 * pow(pyref.MONT_R, -1, R) % R
    assert H.to_int(got) == want
    assert (R - 1) ** 2 * 64 * 300 >= 1 << 512  # the top limbs were in use
    rng = np.random.default_rng(5)
    a = np.stack([H.le32(int.from_bytes(rng.bytes(32), "little") % R) for _ in range(33)])
    b = np.roll(a, 7, axis=0)
    want = sum(H.to_int(x) * H.to_int(y) for x, y in zip(a, b)) * pow(pyref.MONT_R, -1, R) % R
    assert H.to_int(emul.wide_dot(a, b)) == want
    z = np.zeros((4, 32), np.uint8)
    assert H.to_int(emul.wide_dot(z, z)) == 0


def test_ingest_body():
    vals = [0, 1, R - 1, R - 2, 1 << 253] + _rand_fr(3, 20)
    data = np.stack([H.le32(v) for v in vals])
    assert emul.ingest(data) == 0
    assert [pyref.from_mont(H.to_int(row), R) for row in data] == vals
    bad = np.stack([H.le32(5), H.le32(R), H.le32(7)])
    assert emul.ingest(bad) == 1
    assert emul.ingest(np.stack([H.le32((1 << 256) - 1)])) == 1


@pytest.mark.parametrize("nbytes,signed", [(1, False), (2, False), (4, False), (8, False), (8, True)])
def test_widen_body(nbytes, signed):
    rng = np.random.default_rng(nbytes)
    n = 50
    raw = rng.integers(0, 256, size=n * nbytes, dtype=np.uint8)
    raw[:nbytes] = 0xFF  # all-ones: the largest unsigned value / -1
    out = emul.widen(raw, nbytes, signed)
    for i in range(n):
        v = int.from_bytes(bytes(raw[i * nbytes:(i + 1) * nbytes]), "little", signed=signed)
        assert H.to_int(out[i]) == v % R


@pytest.mark.parametrize("party", [0, 1, 2])
def test_lincomb_body(party):
    n = 16
    sh1 = list(zip(_rand_fr(10, n), _rand_fr(11, n)))
    sh2 = list(zip(_rand_fr(12, n // 2), _rand_fr(13, n // 2)))    # a shorter shared polynomial
    pub = _rand_fr(14, n)
    small = [pyref.limb(15, i, 0) & 0xFFFF for i in range(n // 4)]  # short public polynomial with u16 values
    neg = [(-(pyref.limb(16, i, 0) & 0xFFFFFFFF)) % R for i in range(n)]  # i64 negatives as field elements
    coeffs = _rand_fr(17, 5)
    coeffs[1] = 0
    coeffs[3] = 1
    polys_ref = [("shared", sh1), ("shared", sh2), ("public", pub), ("public", small), ("public", neg)]
    kind, want = rep3ref.linear_combination(polys_ref, coeffs, party)
    assert kind == "shared"
    polys = [("shared", _shared_mont(sh1)), ("shared", _shared_mont(sh2)), ("mont", H.scalars_wire(pub)),
             ("canon", H.scalars_wire(small, form=1)), ("canon", H.scalars_wire(neg, form=1))]
    got = emul.lincomb(polys, H.scalars_wire(coeffs), party)
    assert _from_shared(got) == want


@pytest.mark.parametrize("party", [0, 1, 2])
def test_lincomb_split_over_devices(party):
    """The multi-device form: one partial combination per device, then the summing body (the kernel that reads the
    remote partials over NVLink).  Any placement gives the reference's joint polynomial - a device that holds public
    polynomials only contributes a PUBLIC partial, which must follow add_public exactly once."""
    n = 16
    sh1 = list(zip(_rand_fr(10, n), _rand_fr(11, n)))
    sh2 = list(zip(_rand_fr(12, n // 2), _rand_fr(13, n // 2)))
    pub = _rand_fr(14, n)
    small = [pyref.limb(15, i, 0) & 0xFFFF for i in range(n // 4)]
    neg = [(-(pyref.limb(16, i, 0) & 0xFFFFFFFF)) % R for i in range(n)]
    coeffs = _rand_fr(17, 5)
    polys_ref = [("shared", sh1), ("shared", sh2), ("public", pub), ("public", small), ("public", neg)]
    _, want = rep3ref.linear_combination(polys_ref, coeffs, party)
    polys = [("shared", _shared_mont(sh1)), ("shared", _shared_mont(sh2)), ("mont", H.scalars_wire(pub)),
             ("canon", H.scalars_wire(small, form=1)), ("canon", H.scalars_wire(neg, form=1))]
    for placement in ([0, 1, 0, 1, 0], [0, 0, 1, 1, 1], [3, 1, 2, 0, 2], [1, 0, 0, 0, 0]):
        got = emul.lincomb_by_device(polys, H.scalars_wire(coeffs), party, placement)
        assert _from_shared(got) == want, placement
    # public polynomials only: the result stays public whatever the placement
    _, wantp = rep3ref.linear_combination(polys_ref[2:], coeffs[2:], party)
    got = emul.lincomb_by_device(polys[2:], H.scalars_wire(coeffs[2:]), party, [0, 1, 2])
    assert [pyref.from_mont(H.to_int(row), R) for row in got] == wantp


def test_lincomb_public_only_and_degenerate():
    n = 8
    pub = _rand_fr(20, n)
    small = [pyref.limb(21, i, 0) & 0xFF for i in range(n)]
    coeffs = _rand_fr(22, 2)
    kind, want = rep3ref.linear_combination([("public", pub), ("public", small)], coeffs, 1)
    assert kind == "public"
    got = emul.lincomb([("mont", H.scalars_wire(pub)), ("canon", H.scalars_wire(small, form=1))], H.scalars_wire(coeffs), 1)
    assert [pyref.from_mont(H.to_int(row), R) for row in got] == want
    # co-jolt party-0 shape: constant share vectors, gamma powers as coefficients (opening_proof.rs:268-278)
    gamma = pyref.scalar_uniform(23, 0)
    gp = [pow(gamma, j, R) for j in range(6)]
    shares = [[(pyref.scalar_uniform(24, j), pyref.scalar_uniform(25, j))] * n for j in range(6)]
    kind, want = rep3ref.linear_combination([("shared", s) for s in shares], gp, 0)
    got = emul.lincomb([("shared", _shared_mont(s)) for s in shares], H.scalars_wire(gp), 0)
    assert _from_shared(got) == want
    assert len(set(want)) == 1
    with pytest.raises(ValueError):  # the reference panics when a public polynomial outruns every shared one
        rep3ref.linear_combination([("shared", shares[0][:4]), ("public", pub)], gp[:2], 0)


def test_chi_bodies():
    n = 64
    sh = list(zip(_rand_fr(30, n), _rand_fr(31, n)))
    pub = _rand_fr(32, n)
    small = [pyref.limb(33, i, 0) & 0xFFFFFFFF for i in range(n)]
    chis = _rand_fr(34, n)
    chis[0], chis[1] = 0, 1  # mul_public_01_optimized's special cases change nothing
    want = [rep3ref.evaluate_at_chi(("shared", sh), chis), rep3ref.evaluate_at_chi(("public", pub), chis),
            rep3ref.evaluate_at_chi(("public", small), chis)]
    for T in (1, 32, 64):
        got = emul.chi([("shared", _shared_mont(sh)), ("mont", H.scalars_wire(pub)), ("canon", H.scalars_wire(small, form=1))],
                       H.scalars_wire(chis), T=T)
        assert [pyref.from_mont(H.to_int(row), R) for row in got] == want
    # an additive share: the three parties' evaluations add up to the evaluation of the secret
    secret = _rand_fr(35, n)
    t0, t1 = _rand_fr(36, n), _rand_fr(37, n)
    t2 = [(s - x - y) % R for s, x, y in zip(secret, t0, t1)]
    parts = [list(zip(t0, t2)), list(zip(t1, t0)), list(zip(t2, t1))]  # generate_shares_rep3 (mpc-core .. arithmetic.rs:20-32)
    total = sum(rep3ref.evaluate_at_chi(("shared", p), chis) for p in parts) % R
    assert total == sum(s * c for s, c in zip(secret, chis)) % R


def test_pair_sum_body_and_open_identity(orc):
    """S[b] = P[2b] + P[2b+1], with the exceptional pairs; and the identity the paired opening relies on:
    MSM(P, q duplicated) == MSM(S, q)  (pst13.rs:459)."""
    n = 16
    pts = [pyref.base_point(4, i) for i in range(n)]
    pts[3] = pts[2]                 # P + P
    pts[5] = pyref.neg(pts[4])      # P + (-P) = infinity
    bases = H.bases_wire(pts)
    inf = np.zeros(n, np.uint8)
    inf[6] = 1                      # infinity + Q = Q
    inf[8] = inf[9] = 1             # infinity + infinity
    ref_pts = [None if inf[i] else pts[i] for i in range(n)]
    want = rep3ref.pair_sums(ref_pts)
    out, oi = emul.pair_sum(bases, inf)
    for b in range(n // 2):
        if want[b] is None:
            assert oi[b] == 1
        else:
            assert oi[b] == 0 and (out[b] == H.bases_wire([want[b]])[0]).all(), b
    q = _rand_fr(6, n // 2)
    dup = [q[x >> 1] for x in range(n)]
    lhs = pyref.msm_naive(dup, [p if p is not None else None for p in ref_pts]) if all(p is not None for p in ref_pts) else None
    acc = None
    for s, p in zip(dup, ref_pts):
        acc = pyref.add(acc, pyref.mul(s, p) if p is not None else None)
    acc2 = None
    for s, p in zip(q, want):
        acc2 = pyref.add(acc2, pyref.mul(s, p) if p is not None else None)
    assert acc == acc2 and lhs is None


def test_eq_table_bodies_and_spartan_worker_restatement():
    """The two thread bodies behind cozk_eq_evals against the Python restatement, both index orders and every split of the
    index bits; and the restated identities they rest on: sum_b p[b] * chi[b] = p.evaluate(point) (what lets co-spartan's
    per-polynomial `evaluate` run as a dot product), msb-first = lsb-first of the reversed point, and the final value of
    the opening fold = the evaluation of the aggregate."""
    for nv in (1, 4, 7):
        point = _rand_fr(40 + nv, nv)
        pm = np.stack([H.fr_mont(t) for t in point])
        for msb in (False, True):
            want = rep3ref.eq_evals(point, msb_first=msb)
            for lo_bits in sorted({0, nv // 2, nv}):
                got = emul.eq(pm, msb_first=msb, lo_bits=lo_bits)
                assert [pyref.from_mont(H.to_int(row), R) for row in got] == want, (nv, msb, lo_bits)
        assert rep3ref.eq_evals(point, msb_first=True) == rep3ref.eq_evals(point[::-1], msb_first=False)
        poly = _rand_fr(50 + nv, 1 << nv)
        chi = rep3ref.eq_evals(point)
        assert sum(p * c for p, c in zip(poly, chi)) % R == rep3ref.dmle_evaluate(poly, point)
    nv, k, num_comms = 5, 4, 3
    polys = [_rand_fr(60 + j, 1 << nv) for j in range(k)]
    point, eta = _rand_fr(70, nv), _rand_fr(71, 1)[0]
    qs, val, evals = rep3ref.distributed_batch_open_poly_worker(polys, point, eta, num_comms)
    assert [len(q) for q in qs] == [1 << (nv - 1 - i) for i in range(nv)]
    assert val == sum(pow(eta, j, R) * evals[j] for j in range(num_comms)) % R   # evaluation is linear in the polynomial
    assert rep3ref.aggregate_poly(eta, [polys[0][:8], polys[1]])[8:] == [eta * v % R for v in polys[1][8:]]  # zip stops at the shorter
