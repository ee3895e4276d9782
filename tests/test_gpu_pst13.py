"""GPU tier: the reference-facing commitment-scheme operations (include/cozk_pst13.h).  test_combine_commitments follows
the reference's own test, co-jolt/src/poly/commitment/pst13.rs:476-547 (linearity of commit under
combine_commitments); the opening is checked against a direct restatement of open() (pst13.rs:428-474) on the oracle."""
import numpy as np
import pytest

from oracle import pyref
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _levels(orc, nv, seed=1):
    # any fixed G1 points exercise the same arithmetic as the eq-basis SRS
    levels, start = [], 0
    for i in range(nv):
        n = 1 << (nv - i)
        levels.append(orc.gen_bases(seed, n, start=start))
        start += n
    return levels


def test_combine_commitments(cozk, ctx, orc):
    pst = cozk.pst13
    nv = 3
    setup = pst.PST13Setup(ctx, _levels(orc, nv))
    n = 1 << nv
    polys = [[pyref.scalar_uniform(40 + j, i) for i in range(n)] for j in range(3)]
    rho = pyref.scalar_uniform(99, 0)
    comms = pst.batch_commit(setup, [H.scalars_wire(p) for p in polys])
    assert all(c.nv == nv for c in comms)
    # combine_commitments(commit(p_i), rho^i) == commit(sum rho^i p_i)
    joint = [sum(pow(rho, j, H.R) * polys[j][i] for j in range(3)) % H.R for i in range(n)]
    joint_c = pst.commit(setup, H.scalars_wire(joint))
    acc = None
    for j, c in enumerate(comms):
        acc = pyref.add(acc, pyref.mul(pow(rho, j, H.R), orc.wire_to_point(c.g_product)))
    assert orc.wire_to_point(joint_c.g_product) == acc
    setup.release()


def test_batch_commit_rep3_matches_reference_semantics(cozk, ctx, orc):
    pst = cozk.pst13
    nv = 10
    n = 1 << nv
    setup = pst.PST13Setup(ctx, [orc.gen_bases(1, n)])
    bases = orc.gen_bases(1, n)
    # three parties' share-a arrays, built the way generate_poly_shares_rep3 does (dense_mlpoly.rs:553-588):
    # party 0: constant c0, party 1: constant c1, party 2: w - c0 - c1;  AoS {a, b} with b = the next party's a
    w = [pyref.limb(7, i + 2, 0) & 0xFFFFFFFF for i in range(n)]
    a_shares = [pyref.scalars("const", 7, n), [pyref.scalar_uniform(7, 1)] * n, pyref.scalars("wminus", 7, n)]
    assert all((a_shares[0][i] + a_shares[1][i] + a_shares[2][i]) % H.R == w[i] for i in range(n))
    public = [pyref.limb(8, i, 0) & 0xFFFF for i in range(n)]
    results = []
    for party in range(3):
        aos = np.zeros((n, 64), np.uint8)
        aos[:, :32] = H.scalars_wire(a_shares[party])
        aos[:, 32:] = H.scalars_wire(a_shares[(party + 1) % 3])
        out = pst.batch_commit_rep3(setup, [aos, H.scalars_wire(public)], [True, False], commit_to_public=(party == 0),
                                    max_num_bits=[0, 16])
        assert (out[0].g_product == orc.msm(bases, aos)).all()       # stride-64 read of share a
        if party == 0:
            assert (out[1].g_product == orc.msm(bases, H.scalars_wire(public))).all()
        else:
            assert out[1] is None                                     # MaybeShared::Public(None)
        results.append(out[0])
    # the coordinator's sum of the three shares commits to the witness itself
    combined = pst.combine_commitment_shares(results)
    assert (combined.g_product == orc.msm(bases, H.scalars_wire(w))).all()
    # key length error
    with pytest.raises(cozk.CozkError) as e:
        pst.batch_commit(setup, [np.zeros((2 * n, 32), np.uint8)])
    assert e.value.code == cozk.ERR_KEY_LENGTH
    setup.release()


def test_batch_commit_packed_small_scalar_polynomials(cozk, ctx, orc):
    """MultilinearPolynomial::{U8..I64Scalars} (multilinear_polynomial.rs:226-268) through the host boundary in their packed
    form (1 - 8 bytes per coefficient over PCIe, widened on the device): same commitments as the widened 32-byte images and
    the oracle; I64 negatives are full-width field elements; public polynomials are dropped on parties 1 / 2."""
    pst, rep3 = cozk.pst13, cozk.rep3
    nv = 11
    n = 1 << nv
    bases = orc.gen_bases(1, n)
    setup = pst.PST13Setup(ctx, [bases])
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 2, n).astype(np.uint8)                   # 0/1 flags
    u16 = rng.integers(0, 1 << 16, n).astype(np.uint16)
    u32 = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    u64 = rng.integers(0, 1 << 63, n, dtype=np.uint64) * 2 + 1
    i64 = rng.integers(-(1 << 62), 1 << 62, n).astype(np.int64)
    i64[0], i64[1] = -1, np.iinfo(np.int64).min
    large = [pyref.scalar_uniform(21, i) for i in range(n)]
    share = np.zeros((n, 64), np.uint8)
    share[:, :32] = H.scalars_wire([pyref.scalar_uniform(22, i) for i in range(n)])
    share[:, 32:] = H.scalars_wire([pyref.scalar_uniform(23, i) for i in range(n)])
    polys = [u8, u16, share, u32, u64, i64, H.scalars_wire(large)]
    kinds = [rep3.U8, rep3.U16, rep3.SHARED, rep3.U32, rep3.U64, rep3.I64, rep3.PUBLIC]
    ints = [[int(x) for x in u8], [int(x) for x in u16], None, [int(x) for x in u32], [int(x) for x in u64],
            [int(x) % H.R for x in i64], large]
    got = pst.batch_commit_packed(setup, polys, kinds, commit_to_public=True)
    for j, vals in enumerate(ints):
        want = orc.msm(bases, share[:, :32].copy() if vals is None else H.scalars_wire(vals))
        assert got[j].nv == nv and (got[j].g_product == want).all(), j
    other = pst.batch_commit_packed(setup, polys, kinds, commit_to_public=False)      # parties 1 / 2
    assert [c is not None for c in other] == [False, False, True, False, False, False, False]
    assert (other[2].g_product == got[2].g_product).all()
    with pytest.raises(ValueError):
        pst.batch_commit_packed(setup, [u8, u16[:-1]], [rep3.U8, rep3.U16])
    with pytest.raises(cozk.CozkError) as e:                                          # longer than the SRS: Key length error
        pst.batch_commit_packed(setup, [np.zeros(2 * n, np.uint8)], [rep3.U8])
    assert e.value.code == cozk.ERR_KEY_LENGTH
    setup.release()


def _open_reference(orc, levels, evals, point):
    """open() restated on Python ints + the oracle MSM (pst13.rs:428-474)."""
    nv = len(point)
    r = list(evals)
    proofs = []
    for i in range(nv):
        k = nv - i
        t = point[i]
        half = 1 << (k - 1)
        q = [(r[2 * b + 1] - r[2 * b]) % H.R for b in range(half)]
        r = [(r[2 * b] * (1 - t) + r[2 * b + 1] * t) % H.R for b in range(half)]
        scalars = [q[x >> 1] for x in range(1 << k)]
        proofs.append(orc.msm(levels[i], H.scalars_wire(scalars)))
    return np.stack(proofs), r[0]


@pytest.mark.parametrize("nv,stride", [(1, 32), (4, 32), (9, 64)])
def test_open(cozk, ctx, orc, nv, stride):
    pst = cozk.pst13
    levels = _levels(orc, nv, seed=5)
    setup = pst.PST13Setup(ctx, levels)
    n = 1 << nv
    evals = [pyref.scalar_uniform(60, i) for i in range(n)]
    point = [pyref.scalar_uniform(61, i) for i in range(nv)]
    proofs, ev = pst.open(setup, H.scalars_wire(evals, stride=stride), H.scalars_wire(point), stride=stride)
    want_proofs, want_ev = _open_reference(orc, levels, evals, point)
    assert (proofs == want_proofs).all()
    assert pyref.from_mont(H.to_int(ev), H.R) == want_ev
    setup.release()


@pytest.mark.parametrize("nv", [1, 3, 4])
def test_opening_satisfies_the_verifier_equation(cozk, ctx, orc, nv):
    """The reference accepts an opening through the pairing check (pst13.rs:536-545: PST13::verify -> MultilinearPC::check).
    With a real eq-basis SRS (oracle/pairing.py: powers_of_g from a seeded trapdoor, h and h_mask in G2) the engine's
    commitment, proofs and evaluation - host path, keyed path and resident-polynomial path - satisfy
    e(C - v g, h) = prod e(proof_i, h_mask_i - point_i h); the same proofs fail it for a wrong value or a reversed point.
    Unlike _open_reference this witness shares no code with open(): a wrong level / point / base order would not pass."""
    from oracle import pairing as pr
    pst, rep3 = cozk.pst13, cozk.rep3
    levels_pts, vk, _ = pr.setup(nv, seed=70 + nv)
    levels = [H.bases_wire(lv) for lv in levels_pts]
    setup = rep3.create_open_key(pst.PST13Setup(ctx, levels))
    n = 1 << nv
    evals = [pyref.scalar_uniform(80 + nv, i) for i in range(n)]
    point = [pyref.scalar_uniform(90 + nv, i) for i in range(nv)]
    comm = orc.wire_to_point(pst.commit(setup, H.scalars_wire(evals)).g_product)
    poly = rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(evals), rep3.PUBLIC)
    runs = {"host": pst.open(setup, H.scalars_wire(evals), H.scalars_wire(point)),
            "keyed": rep3.open_keyed(setup, H.scalars_wire(evals), H.scalars_wire(point)),
            "resident": rep3.open_poly(setup, poly, H.scalars_wire(point)),
            "resident, reference schedule": rep3.open_poly(setup, poly, H.scalars_wire(point), keyed=False)}
    for name, (proofs, ev) in runs.items():
        value = pyref.from_mont(H.to_int(ev), H.R)
        pts = [orc.wire_to_point(p) for p in proofs]
        assert pr.verify_opening(vk, comm, point, value, pts), name
    proofs, ev = runs["host"]
    pts = [orc.wire_to_point(p) for p in proofs]
    value = pyref.from_mont(H.to_int(ev), H.R)
    assert not pr.verify_opening(vk, comm, point, (value + 1) % H.R, pts)
    if nv > 1:
        assert not pr.verify_opening(vk, comm, point[::-1], value, pts)
    poly.release()
    rep3.release_open_key(setup)
    setup.release()


def test_fixed_base_batch_mul_srs_generation(cozk, ctx, orc):
    """SRS generation, G1 half: out[i] = s_i * g against the oracle's double-and-add, including
    s = 0, 1, r - 1 and a non-generator base; the registered handle commits like an uploaded SRS."""
    n = 600
    scal = [0, 1, 2, H.R - 1, H.R - 2, 255, 256, 1 << 200] + [pyref.scalar_uniform(77, i) for i in range(n - 8)]
    for base_pt in (pyref.G1, pyref.base_point(3, 5)):
        base = H.point_wire(base_pt)
        for form in (0, 1):
            pts, _ = ctx.fixed_base_batch_mul(base, H.scalars_wire(scal, form), form=form)
            for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 99, n - 1):
                want = orc.g1_mul(base, H.le32(scal[i]), form=orc.CANON)
                assert (pts[i] == want).all(), (i, form)
    # all points at once through linearity: sum_i s_i * g = (sum_i s_i) * g
    base = H.point_wire(pyref.G1)
    pts, srs = ctx.fixed_base_batch_mul(base, H.scalars_wire(scal), register=True)
    total = cozk.g1_sum(pts)
    assert (total == orc.g1_mul(base, H.le32(sum(scal) % H.R), form=orc.CANON)).all()
    # the registered SRS (with its infinity flag for s = 0) commits like the same points uploaded by hand
    assert ctx.srs_len(srs) == n
    sc = orc.gen_scalars("uniform", 12, n)
    manual = ctx.srs_register(pts, infinity=pts[:, 64].copy(), stride=72)
    assert (ctx.msm_batch(srs, sc)[0] == ctx.msm_batch(manual, sc)[0]).all()
    masked = sc.copy()
    masked[0] = 0
    bases = pts[:, :64].copy()
    bases[0] = orc.gen_bases(1, 1)[0]
    assert (ctx.msm_batch(srs, sc)[0] == orc.msm(bases, masked)).all()
    ctx.srs_release(srs)
    ctx.srs_release(manual)
    ident, _ = ctx.fixed_base_batch_mul(H.point_wire(None), H.scalars_wire(scal[:5]))
    assert (ident[:, 64] == 1).all()
