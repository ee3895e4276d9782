"""GPU tier for SURVEY.md 8(f) rows N1 / N2 / N4 through the C ABI (include/cozk_rep3.h): wire ingestion, commitment of
resident polynomials, linear combination, chi dot products, pair-sum SRS and the opening of a resident polynomial -
bit-exact against oracle/rep3ref.py (Python integers), the C oracle's MSM, and the reference's own linearity property
(co-jolt/src/poly/commitment/pst13.rs:498-546) at sizes the Python oracle cannot reach."""
import numpy as np
import pytest

from oracle import pyref, rep3ref
from tests import helpers as H
from tests.test_gpu_pst13 import _levels, _open_reference

pytestmark = pytest.mark.gpu
R = H.R


def _rand_fr(seed, n):
    return [pyref.scalar_uniform(seed, i) for i in range(n)]


def _shared_mont(vals):
    out = np.zeros((len(vals), 64), np.uint8)
    out[:, :32] = H.scalars_wire([a for a, _ in vals])
    out[:, 32:] = H.scalars_wire([b for _, b in vals])
    return out


def _from_shared(arr):
    return [(pyref.from_mont(H.to_int(row[:32]), R), pyref.from_mont(H.to_int(row[32:]), R)) for row in arr]


def _from_dense(arr):
    return [pyref.from_mont(H.to_int(row), R) for row in arr]


def test_from_wire(cozk, ctx):
    rep3 = cozk.rep3
    n = 256
    coeffs = list(zip(_rand_fr(1, n), _rand_fr(2, n)))
    coeffs[0] = (0, R - 1)
    for tagged in (False, True):
        raw = rep3ref.serialize_rep3_dense(coeffs, tagged=tagged)
        poly, used = rep3.Rep3DensePolynomial.from_wire(ctx, raw + b"next message", tagged=tagged)
        assert used == len(raw) and len(poly) == n and poly.kind == rep3.SHARED
        assert (poly.download() == _shared_mont(coeffs)).all()
        poly.release()
    # chunk_range (split_poly, dense_mlpoly.rs:266-292), bound coefficients and scratch space present
    raw = rep3ref.serialize_rep3_dense(coeffs, bound_coeffs=coeffs[:3], scratch=coeffs[:5], chunk_range=(64, 128))
    poly, used = rep3.Rep3DensePolynomial.from_wire(ctx, raw)
    assert used == len(raw) and len(poly) == 64
    assert (poly.download() == _shared_mont(coeffs[64:128])).all()
    poly.release()
    # the errors ark-serialize reports
    for bad in (raw[:-1], raw[:100], b""):
        with pytest.raises(cozk.CozkError) as e:
            rep3.Rep3DensePolynomial.from_wire(ctx, bad)
        assert e.value.code == cozk.ERR_WIRE
    with pytest.raises(cozk.CozkError) as e:
        rep3.Rep3DensePolynomial.from_wire(ctx, b"\x00" + raw, tagged=True)  # Public variant
    assert e.value.code == cozk.ERR_WIRE
    over = bytearray(rep3ref.serialize_rep3_dense(coeffs[:4]))
    over[16 + 64:16 + 96] = int(R).to_bytes(32, "little")  # coefficient 1, share a = r
    with pytest.raises(cozk.CozkError) as e:
        rep3.Rep3DensePolynomial.from_wire(ctx, bytes(over))
    assert e.value.code == cozk.ERR_WIRE
    empty, _ = rep3.Rep3DensePolynomial.from_wire(ctx, rep3ref.serialize_rep3_dense([]))
    assert len(empty) == 0
    empty.release()


def test_commit_resident_polynomials(cozk, ctx, orc):
    """PST13::batch_commit_rep3 over handles: share a read at stride 64, public and small-scalar polynomials on party 0."""
    rep3, pst = cozk.rep3, cozk.pst13
    nv = 11
    n = 1 << nv
    bases = orc.gen_bases(1, n)
    setup = pst.PST13Setup(ctx, [bases])
    a, b = _rand_fr(3, n), _rand_fr(4, n)
    pub = _rand_fr(5, n)
    rng = np.random.default_rng(6)
    smalls = {rep3.U8: rng.integers(0, 1 << 8, n), rep3.U16: rng.integers(0, 1 << 16, n), rep3.U32: rng.integers(0, 1 << 32, n),
              rep3.U64: rng.integers(0, 1 << 63, n, dtype=np.uint64) * 2 + 1, rep3.I64: rng.integers(-(1 << 62), 1 << 62, n)}
    polys = [rep3.Rep3DensePolynomial.from_wire(ctx, rep3ref.serialize_rep3_dense(list(zip(a, b))))[0],
             rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(pub), rep3.PUBLIC)]
    polys += [rep3.Rep3DensePolynomial.upload(ctx, v, k) for k, v in smalls.items()]
    want = [orc.msm(bases, H.scalars_wire(a)), orc.msm(bases, H.scalars_wire(pub))]
    want += [orc.msm(bases, H.scalars_wire([int(x) % R for x in v])) for v in smalls.values()]
    out = rep3.batch_commit_rep3(setup, polys, commit_to_public=True)
    for j, c in enumerate(out):
        assert c.nv == nv and (c.g_product == want[j]).all(), j
    out = rep3.batch_commit_rep3(setup, polys, commit_to_public=False)
    assert (out[0].g_product == want[0]).all() and all(c is None for c in out[1:])
    # small kinds come back as canonical integers
    assert [H.to_int(r) for r in polys[6].download()] == [int(x) % R for x in smalls[rep3.I64]]
    short = rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(pub[:n // 2]), rep3.PUBLIC)
    with pytest.raises(cozk.CozkError):  # "batch commit requires all batches to have the same length"
        rep3.batch_commit_rep3(setup, [polys[0], short], True)
    for p in polys + [short]:
        p.release()
    setup.release()


@pytest.mark.parametrize("party", [0, 1, 2])
def test_linear_combination_vs_python(cozk, ctx, party):
    rep3 = cozk.rep3
    n = 96
    sh1 = list(zip(_rand_fr(10, n), _rand_fr(11, n)))
    sh2 = list(zip(_rand_fr(12, n // 3), _rand_fr(13, n // 3)))
    pub = _rand_fr(14, n)
    u16 = [pyref.limb(15, i, 0) & 0xFFFF for i in range(n // 2)]
    i64 = [(pyref.limb(16, i, 0) & 0xFFFFFFFF) - (1 << 31) for i in range(n)]
    coeffs = _rand_fr(17, 5)
    coeffs[1], coeffs[3] = 1, 0
    ref = [("shared", sh1), ("shared", sh2), ("public", pub), ("public", u16), ("public", [v % R for v in i64])]
    _, want = rep3ref.linear_combination(ref, coeffs, party)
    polys = [rep3.Rep3DensePolynomial.upload(ctx, _shared_mont(sh1)), rep3.Rep3DensePolynomial.upload(ctx, _shared_mont(sh2)),
             rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(pub), rep3.PUBLIC),
             rep3.Rep3DensePolynomial.upload(ctx, u16, rep3.U16), rep3.Rep3DensePolynomial.upload(ctx, i64, rep3.I64)]
    joint = rep3.linear_combination(polys, H.scalars_wire(coeffs), party)
    assert joint.kind == rep3.SHARED and len(joint) == n
    assert _from_shared(joint.download()) == want
    # all-public inputs stay public, whatever the party
    _, wantp = rep3ref.linear_combination(ref[2:], coeffs[2:], party)
    jp = rep3.linear_combination(polys[2:], H.scalars_wire(coeffs[2:]), party)
    assert jp.kind == rep3.PUBLIC and _from_dense(jp.download()) == wantp
    # a public polynomial that outruns every shared one: the reference panics in as_shared()
    with pytest.raises(cozk.CozkError) as e:
        rep3.linear_combination([polys[1], polys[2]], H.scalars_wire(coeffs[:2]), party)
    assert e.value.code == cozk.ERR_INVALID_ARG
    for p in polys + [joint, jp]:
        p.release()


def test_linear_combination_commutes_with_commit(cozk, ctx, orc):
    """The reference's own property (pst13.rs:498-546) at a size Python integers cannot reach:
    commit(sum_j gamma^j p_j) == sum_j gamma^j commit(p_j), share a of every party shape."""
    rep3, pst = cozk.rep3, cozk.pst13
    nv, k = 14, 12
    n = 1 << nv
    bases = orc.gen_bases(1, n)
    setup = pst.PST13Setup(ctx, [bases])
    gamma = pyref.scalar_uniform(40, 0)
    gp = [pow(gamma, j, R) for j in range(k)]
    polys = []
    for j in range(k):
        aos = np.zeros((n, 64), np.uint8)
        aos[:, :32] = orc.gen_scalars(("uniform", "const", "wminus")[j % 3], 50 + j, n)
        aos[:, 32:] = orc.gen_scalars("uniform", 90 + j, n)
        polys.append(rep3.Rep3DensePolynomial.upload(ctx, aos))
    comms = rep3.batch_commit_rep3(setup, polys, False)
    joint = rep3.linear_combination(polys, H.scalars_wire(gp), 2)
    cj = rep3.batch_commit_rep3(setup, [joint], False)[0]
    acc = None
    for j, c in enumerate(comms):
        acc = pyref.add(acc, pyref.mul(gp[j], orc.wire_to_point(c.g_product)))
    assert orc.wire_to_point(cj.g_product) == acc
    # and element-wise against the C oracle's vectorised field arithmetic, both halves
    dl = joint.download()
    for half in (0, 1):
        want = np.zeros((n, 4), np.uint64)
        for j, p in enumerate(polys):
            col = np.ascontiguousarray(p.download()[:, 32 * half:32 * half + 32]).view(np.uint64)
            cj_ = np.tile(H.fr_mont(gp[j]).view(np.uint64), (n, 1))
            want = orc.field_op("fr", "add", want, orc.field_op("fr", "mul", col, cj_))
        assert (np.ascontiguousarray(dl[:, 32 * half:32 * half + 32]).view(np.uint64) == want).all(), half
    for p in polys + [joint]:
        p.release()
    setup.release()


def test_evaluate_at_chi(cozk, ctx):
    rep3 = cozk.rep3
    for n in (1, 32, 1000):
        sh = list(zip(_rand_fr(30, n), _rand_fr(31, n)))
        pub = _rand_fr(32, n)
        u32 = [pyref.limb(33, i, 0) & 0xFFFFFFFF for i in range(n)]
        chis = _rand_fr(34, n)
        polys = [rep3.Rep3DensePolynomial.upload(ctx, _shared_mont(sh)),
                 rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(pub), rep3.PUBLIC),
                 rep3.Rep3DensePolynomial.upload(ctx, u32, rep3.U32)]
        got = rep3.batch_evaluate_at_chi(polys, H.scalars_wire(chis))
        want = [rep3ref.evaluate_at_chi(("shared", sh), chis), rep3ref.evaluate_at_chi(("public", pub), chis),
                rep3ref.evaluate_at_chi(("public", u32), chis)]
        assert _from_dense(got) == want, n
        assert _from_dense(polys[0].evaluate_at_chi(H.scalars_wire(chis))[None])[0] == want[0]
        with pytest.raises(cozk.CozkError):  # zip_eq
            rep3.batch_evaluate_at_chi(polys, H.scalars_wire(chis + [1]))
        for p in polys:
            p.release()


def test_pair_sum_srs(cozk, ctx, orc):
    """cozk_srs_pair_sums: MSM(P, q duplicated) == MSM(S, q), exceptional pairs included (pst13.rs:459)."""
    rep3 = cozk.rep3
    n = 2048
    bases = orc.gen_bases(3, n)
    bases[3] = bases[2]
    bases[5, :32] = bases[4, :32]
    bases[5, 32:] = H.fq_mont((-pyref.from_mont(H.to_int(bases[4, 32:]), H.P)) % H.P)
    srs = ctx.srs_register(bases)
    pairs = rep3.pair_sums(ctx, srs)
    assert ctx.srs_len(pairs) == n // 2
    q = orc.gen_scalars("uniform", 8, n // 2)
    assert (ctx.msm_batch(pairs, q)[0] == orc.msm(bases, np.repeat(q, 2, axis=0))).all()
    ctx.srs_release(pairs)
    ctx.srs_release(srs)


@pytest.mark.parametrize("nv,small", [(1, 12), (5, 12), (10, 12), (10, 3), (10, 0), (15, 12), (15, 11), (15, 6)])
def test_open_resident_polynomial_with_key(cozk, ctx, orc, nv, small):
    """open() on a resident polynomial, with the opening key (pair sums, batched small levels) and with the reference's
    schedule, against the restated open()."""
    rep3, pst = cozk.rep3, cozk.pst13
    levels = _levels(orc, nv, seed=7)
    if nv >= 5:
        levels[0][3] = levels[0][2]  # P + P inside a pair
        levels[1][1, 32:] = H.fq_mont((-pyref.from_mont(H.to_int(levels[1][0, 32:]), H.P)) % H.P)
        levels[1][1, :32] = levels[1][0, :32]  # P + (-P): the pair sum is the point at infinity
    ctx.set_option("open_small_log2", small)
    ctx.set_option("open_one_batch_max_nv", 0 if small != 12 else 20)  # 0: split the levels by open_small_log2 as asked
    try:
        setup = rep3.create_open_key(pst.PST13Setup(ctx, levels))
    finally:
        ctx.set_option("open_small_log2", 15)
        ctx.set_option("open_one_batch_max_nv", 20)
    n = 1 << nv
    a, b = _rand_fr(60, n), _rand_fr(62, n)
    point = _rand_fr(61, nv)
    want_proofs, want_ev = _open_reference(orc, levels, a, point)
    poly = rep3.Rep3DensePolynomial.upload(ctx, _shared_mont(list(zip(a, b))))
    for keyed in (True, False):
        proofs, ev = rep3.open_poly(setup, poly, H.scalars_wire(point), keyed=keyed)
        assert (proofs == want_proofs).all(), keyed
        assert pyref.from_mont(H.to_int(ev), R) == want_ev
    # the small levels as a zero-padded batch instead of a ragged one: the same proofs
    ctx.set_option("open_small_ragged", 0)
    try:
        proofs, _ = rep3.open_poly(setup, poly, H.scalars_wire(point), keyed=True)
        assert (proofs == want_proofs).all()
    finally:
        ctx.set_option("open_small_ragged", 1)
    proofs, ev = rep3.open_keyed(setup, H.scalars_wire(a), H.scalars_wire(point))
    assert (proofs == want_proofs).all() and pyref.from_mont(H.to_int(ev), R) == want_ev
    # prove_rep3 reverses the opening point first (pst13.rs:134)
    proofs, _ = rep3.prove_rep3(setup, poly, H.scalars_wire(point[::-1]))
    assert (proofs == want_proofs).all()
    # a public polynomial opens too (open() itself takes a DensePolynomial)
    dense = rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(a), rep3.PUBLIC)
    proofs, _ = rep3.open_poly(setup, dense, H.scalars_wire(point))
    assert (proofs == want_proofs).all()
    wrong = rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(a + a), rep3.PUBLIC)
    with pytest.raises(cozk.CozkError) as e:  # assert_eq!(nv, ck.nv, "Invalid size of polynomial")
        rep3.open_poly(setup, wrong, H.scalars_wire(point))
    assert e.value.code == cozk.ERR_KEY_LENGTH
    for p in (poly, dense, wrong):
        p.release()
    rep3.release_open_key(setup)
    setup.release()


def test_open_key_lifecycle_is_deterministic(cozk, ctx):
    """Keys, pair-sum SRSs and polynomials created and released in a loop (pooled, stream-ordered allocations): the keyed
    and the reference schedule must agree every time, and with the first round."""
    rep3 = cozk.rep3
    nv = 17
    handles, start = [], 0
    for i in range(nv):
        m = 1 << (nv - i)
        d = ctx.testgen_bases(3, m, start=start)
        handles.append(ctx.srs_register_device(d, m))
        d.free()
        start += m

    class Setup:
        pass
    setup = Setup()
    setup.ctx, setup.level_srs = ctx, handles
    point = ctx.testgen_scalars("uniform", 12, nv).download().reshape(nv, 32)
    first = None
    for rnd in range(4):
        tmp = ctx.testgen_scalars("uniform", 11, 1 << nv, stride=64)
        poly = rep3.Rep3DensePolynomial.from_device(ctx, tmp, 1 << nv)
        tmp.free()
        rep3.create_open_key(setup)
        keyed, _ = rep3.open_poly(setup, poly, point, keyed=True)
        plain, _ = rep3.open_poly(setup, poly, point, keyed=False)
        assert (keyed == plain).all(), rnd
        first = keyed if first is None else first
        assert (keyed == first).all(), rnd
        rep3.release_open_key(setup)
        poly.release()
    for h in handles:
        ctx.srs_release(h)


def test_three_party_flow_commit_combine_open(cozk, ctx, orc):
    """The reference's flow end to end on resident shares (witness.rs:303-365, opening_proof.rs:255-288, pst13.rs:72-137):
    every party ingests its wire share, commits, forms the joint polynomial and opens it; the coordinator's sums equal
    the commitment / opening proof of the plain witness polynomials."""
    rep3, pst = cozk.rep3, cozk.pst13
    nv, k = 6, 3
    n = 1 << nv
    levels = _levels(orc, nv, seed=9)
    setup = rep3.create_open_key(pst.PST13Setup(ctx, levels))
    secrets = [[pyref.limb(70 + j, i, 0) & 0xFFFFFFFF for i in range(n)] for j in range(k)]
    public = [pyref.limb(80, i, 0) & 0xFF for i in range(n)]
    gamma = pyref.scalar_uniform(81, 0)
    gp = [pow(gamma, j, R) for j in range(k + 1)]
    point = _rand_fr(82, nv)
    party_comms, party_proofs = [], []
    for party in range(3):
        polys = []
        for j in range(k):
            t = [_rand_fr(100 + 10 * j, n), _rand_fr(101 + 10 * j, n)]
            t.append([(s - x - y) % R for s, x, y in zip(secrets[j], t[0], t[1])])
            share = list(zip(t[party], t[(party + 2) % 3]))  # party i holds (t_i, t_{i-1}): arithmetic.rs:20-32
            polys.append(rep3.Rep3DensePolynomial.from_wire(ctx, rep3ref.serialize_rep3_dense(share, tagged=True), tagged=True)[0])
        polys.append(rep3.Rep3DensePolynomial.upload(ctx, public, rep3.U8))
        comms = rep3.batch_commit_rep3(setup, polys, commit_to_public=(party == 0))
        assert (comms[k] is not None) == (party == 0)
        party_comms.append(comms)
        joint = rep3.linear_combination(polys, H.scalars_wire(gp), party)
        proofs, _ = rep3.prove_rep3(setup, joint, H.scalars_wire(point[::-1]))
        party_proofs.append(proofs)
        for p in polys + [joint]:
            p.release()
    for j in range(k):
        combined = pst.combine_commitment_shares([party_comms[p][j] for p in range(3)])
        assert (combined.g_product == orc.msm(levels[0], H.scalars_wire(secrets[j]))).all()
    assert (party_comms[0][k].g_product == orc.msm(levels[0], H.scalars_wire(public))).all()
    plain_joint = [(sum(gp[j] * secrets[j][i] for j in range(k)) + gp[k] * public[i]) % R for i in range(n)]
    want_proofs, _ = _open_reference(orc, levels, plain_joint, point)
    assert (pst.coordinate_prove(party_proofs) == want_proofs).all()
    rep3.release_open_key(setup)
    setup.release()


@pytest.mark.parametrize("nv", [1, 2, 9, 14])
def test_eq_evals_on_device(cozk, ctx, nv):
    """cozk_eq_evals against the restated eq table, both index orders; the table then serves evaluate_at_chi from HBM
    (no H2D of the chis) and reproduces DenseMultilinearExtension::evaluate as a dot product."""
    rep3 = cozk.rep3
    point = _rand_fr(80 + nv, nv)
    n = 1 << nv
    for msb in (False, True):
        chi = rep3.eq_evals(ctx, H.scalars_wire(point), msb_first=msb)
        assert len(chi) == n and chi.kind == rep3.PUBLIC
        assert _from_dense(chi.download()) == rep3ref.eq_evals(point, msb_first=msb), (nv, msb)
        chi.release()
    a, b, pub = _rand_fr(81, n), _rand_fr(82, n), _rand_fr(83, n)
    polys = [rep3.Rep3DensePolynomial.upload(ctx, _shared_mont(list(zip(a, b)))),
             rep3.Rep3DensePolynomial.upload(ctx, H.scalars_wire(pub), rep3.PUBLIC)]
    chi = rep3.eq_evals(ctx, H.scalars_wire(point), msb_first=False)
    got = _from_dense(rep3.batch_evaluate_at_chi(polys, chi))
    assert got[1] == rep3ref.dmle_evaluate(pub, point)
    assert got[0] == rep3ref.evaluate_at_chi(("shared", list(zip(a, b))), rep3ref.eq_evals(point))
    # the host-table entry point gives the same bytes
    assert got == _from_dense(rep3.batch_evaluate_at_chi(polys, H.scalars_wire(rep3ref.eq_evals(point))))
    short = rep3.eq_evals(ctx, H.scalars_wire(point[:-1] if nv > 1 else point + point))
    with pytest.raises(cozk.CozkError):  # zip_eq: lengths differ
        rep3.batch_evaluate_at_chi(polys, short)
    with pytest.raises(cozk.CozkError):  # the table must be a polynomial of field elements
        rep3.batch_evaluate_at_chi(polys, polys[0])
    for p in polys + [chi, short]:
        p.release()


@pytest.mark.parametrize("nv,k,num_comms", [(3, 2, 2), (10, 5, 3), (13, 4, 4)])
def test_spartan_batch_open_worker(cozk, ctx, orc, nv, k, num_comms):
    """co-spartan's distributed_batch_open_poly_worker (worker.rs:745-772) in one call over resident polynomials, against
    its restatement: proofs[i] = MSM(ck.powers_of_g[i], q_i duplicated), val = the aggregate at the point, evals = every
    polynomial at the point; keyed and reference opening schedules; and the pieces under the reference's names."""
    rep3, pst, spartan = cozk.rep3, cozk.pst13, cozk.spartan
    levels = _levels(orc, nv, seed=11)
    ck = rep3.create_open_key(pst.PST13Setup(ctx, levels))
    n = 1 << nv
    evals = [_rand_fr(90 + j, n) for j in range(k)]
    point, eta = _rand_fr(97, nv), _rand_fr(98, 1)[0]
    qs, want_val, want_evals = rep3ref.distributed_batch_open_poly_worker(evals, point, eta, num_comms)
    want_proofs = np.stack([orc.msm(levels[i], np.repeat(H.scalars_wire(qs[i]), 2, axis=0)) for i in range(nv)])
    polys = [spartan.upload_evaluations(ctx, H.scalars_wire(e)) for e in evals]
    for keyed in (True, False):
        pf = spartan.distributed_batch_open_poly_worker(polys, ck, H.scalars_wire(point), H.scalars_wire([eta])[0], num_comms, keyed=keyed)
        assert (pf["proofs"] == want_proofs).all(), keyed
        assert pyref.from_mont(H.to_int(pf["val"]), R) == want_val
        assert _from_dense(pf["evals"]) == want_evals
    # the same from the pieces
    agg = spartan.aggregate_poly(H.scalars_wire([eta])[0], polys[:num_comms])
    assert _from_dense(agg.download()) == rep3ref.aggregate_poly(eta, evals[:num_comms])
    proofs, val = spartan.distributed_open(ck, agg, H.scalars_wire(point))
    assert (proofs == want_proofs).all() and pyref.from_mont(H.to_int(val), R) == want_val
    comms = spartan.poly_commit_worker(ck, polys)
    for c, e in zip(comms, evals):
        assert c.nv == nv and (c.g_product == orc.msm(levels[0], H.scalars_wire(e))).all()
    # polys[0..num_comms] out of range panics in the reference
    with pytest.raises(cozk.CozkError):
        spartan.distributed_batch_open_poly_worker(polys, ck, H.scalars_wire(point), H.scalars_wire([eta])[0], k + 1)
    # a polynomial of another size: evaluate() asserts the point length
    odd = spartan.upload_evaluations(ctx, H.scalars_wire(evals[0] + evals[0]))
    with pytest.raises(cozk.CozkError):
        spartan.distributed_batch_open_poly_worker(polys + [odd], ck, H.scalars_wire(point), H.scalars_wire([eta])[0], num_comms)
    for p in polys + [agg, odd]:
        p.release()
    rep3.release_open_key(ck)
    ck.release()


def test_multi_device_resident_flow(cozk, orc):
    """A party's polynomials dealt over several GPUs of one process: every device commits its own (side by side), the
    linear combination is formed per device and summed on device 0 by the kernel that reads the remote partials over
    peer mappings (and by the staged-copy fallback), and the joint polynomial opens there.  Same bytes as the
    one-device flow and the restatement.  On a single-GPU box the context is opened over the same GPU three times
    (helpers.multi_device_ids): every path but the direct peer read then still runs."""
    ids = H.multi_device_ids(4)
    ndev = len(ids)
    rep3, pst = cozk.rep3, cozk.pst13
    nv = 10
    n = 1 << nv
    with cozk.Context(devices=ids) as mctx:
        levels = _levels(orc, nv, seed=13)
        setup = rep3.create_open_key(pst.PST13Setup(mctx, levels))
        shared = [list(zip(_rand_fr(100 + 2 * j, n), _rand_fr(101 + 2 * j, n))) for j in range(5)]
        shared[4] = shared[4][: n // 2]                       # a shorter shared polynomial
        pub = _rand_fr(120, n)
        small = [pyref.limb(121, i, 0) & 0xFFFF for i in range(n)]
        coeffs = _rand_fr(122, 7)
        ref_polys = [("shared", s_) for s_ in shared] + [("public", pub), ("public", small)]

        def upload(placement):
            out = [rep3.Rep3DensePolynomial.upload(mctx, _shared_mont(s_), device=placement[j]) for j, s_ in enumerate(shared)]
            out.append(rep3.Rep3DensePolynomial.upload(mctx, H.scalars_wire(pub), rep3.PUBLIC, device=placement[5]))
            out.append(rep3.Rep3DensePolynomial.upload(mctx, np.array(small, np.uint16), rep3.U16, device=placement[6]))
            return out

        last = ndev - 1
        placements = [[j % ndev for j in range(7)],           # round-robin
                      [0, 0, 0, 0, 0, last, last],            # a device that holds public polynomials only
                      [last] * 7]                             # all on one device: the result stays there
        point = _rand_fr(123, nv)
        for party in (0, 1, 2):
            _, want = rep3ref.linear_combination(ref_polys, coeffs, party)
            for placement in placements:
                polys = upload(placement)
                for direct in (1, 0):
                    mctx.set_option("peer_direct", direct)
                    joint = rep3.linear_combination(polys, H.scalars_wire(coeffs), party)
                    assert joint.info() == (n, rep3.SHARED, 0 if len(set(placement)) > 1 else placement[0])
                    assert _from_shared(joint.download()) == want, (party, placement, direct)
                    if party == 0 and direct == 1 and joint.info()[2] == 0:
                        want_proofs, want_ev = _open_reference(orc, levels, [a for a, _ in want], point)
                        proofs, ev = rep3.open_poly(setup, joint, H.scalars_wire(point))
                        assert (proofs == want_proofs).all() and pyref.from_mont(H.to_int(ev), R) == want_ev
                    joint.release()
                mctx.set_option("peer_direct", 1)
                if party == 0:
                    # commitments of full-length polynomials spread over the devices
                    full = [polys[j] for j in (0, 1, 2, 3, 5, 6)]
                    comms = rep3.batch_commit_rep3(setup, full, commit_to_public=True)
                    vecs = [H.scalars_wire([a for a, _ in shared[j]]) for j in range(4)] + [H.scalars_wire(pub), H.scalars_wire(small)]
                    for c, v in zip(comms, vecs):
                        assert (c.g_product == orc.msm(levels[0], v)).all()
                for p in polys:
                    p.release()
        # public polynomials only, on different devices: the result is public
        pp = [rep3.Rep3DensePolynomial.upload(mctx, H.scalars_wire(pub), rep3.PUBLIC, device=0),
              rep3.Rep3DensePolynomial.upload(mctx, np.array(small, np.uint16), rep3.U16, device=last)]
        joint = rep3.linear_combination(pp, H.scalars_wire(coeffs[:2]), 1)
        _, wantp = rep3ref.linear_combination(ref_polys[5:], coeffs[:2], 1)
        assert joint.kind == rep3.PUBLIC and _from_dense(joint.download()) == wantp
        st = rep3.last_stats(mctx)
        assert st["peer_bytes"] == 32 * n and st["partial_sum_ms"] > 0
        for p in pp + [joint]:
            p.release()
        rep3.release_open_key(setup)
        setup.release()
