"""Python big-integer restatement of the Rep3 polynomial steps either side of the MSM.

TEST INFRASTRUCTURE ONLY (see oracle/pyref.py): imported by tests/ and __graft_entry__.smoke(), never by the product.

PARITY UNPINNED, as for the MSM: the reference ships no fixture for these functions and cannot be built here.  What is
restated, each from the in-tree source it follows:

  wire image            derived CanonicalSerialize of Rep3DensePolynomial, co-jolt/src/poly/dense_mlpoly.rs:23-32:
                        num_vars u64 | coeffs (u64 len | len x (a | b)) | bound_coeffs (same) | Option<scratch> (u8 tag
                        [| vec]) | len u64 | chunk_range (u64, u64); Rep3PrimeFieldShare{a, b}
                        (mpc-types/src/protocols/rep3/arithmetic/types.rs:22-29); every Fr a 32-byte little-endian
                        CANONICAL integer (ark-serialize uncompressed; ark-ff is a third-party crate - the format is
                        its published one); Rep3MultilinearPolynomial prepends a discriminant byte, 1 = Shared
                        (co-jolt/src/poly/multilinear_polynomial.rs:800-821).
  linear_combination    co-jolt/src/poly/multilinear_polynomial.rs:196-296 (per index: sum of coeff * eval over the
                        polynomials long enough to reach it; shared terms via mul_public on both halves; public terms via
                        add_public: share a on party 0, share b on party 1 - TACEO mpc-core, third party, as restated
                        in SURVEY.md 8(f)); the all-public case stays public.
  evaluate_at_chi       co-jolt/src/poly/dense_mlpoly.rs:160-181 with into_additive
                        (mpc-types/src/protocols/rep3/arithmetic/types.rs:76-81): (a + b) * TWO_INV, TWO_INV = (r+1)/2
                        (snarks-core/src/field.rs:6).
  pair sums / open      co-jolt/src/poly/commitment/pst13.rs:428-474; scalars[x] = q[x >> 1] (:459).
  eq tables             chi[b] = prod_i (bit_i(b) ? t_i : 1 - t_i).  ark-poly's DenseMultilinearExtension::evaluate (third
                        party) fixes variable i = bit i of the index with point[i] - the same recursion the in-tree
                        distributed_open writes out (co-noir-spartan/co-spartan/src/worker.rs:793-798:
                        r[k-1][b] = r[k][2b] (1 - point_i) + r[k][2b+1] point_i), which is what dmle_evaluate follows;
                        jolt-core's EqPolynomial::evals (third party) pairs point[0] with the TOP index bit.
  co-spartan worker     aggregate_poly co-noir-spartan/co-spartan/src/utils.rs:85-107; distributed_batch_open_poly_worker
                        co-noir-spartan/co-spartan/src/worker.rs:745-772 (aggregate the first num_comms polynomials, open
                        the aggregate, evaluate EVERY polynomial at the point).
"""
import struct

from . import pyref

R = pyref.R_ORDER
TWO_INV = (R + 1) // 2


# ---------------------------------------------------------------- wire image (ark-serialize, uncompressed)

def _fr(x):
    return int(x % R).to_bytes(32, "little")


def _shares(vec):
    return struct.pack("<Q", len(vec)) + b"".join(_fr(a) + _fr(b) for a, b in vec)


def serialize_rep3_dense(coeffs, bound_coeffs=(), scratch=None, chunk_range=None, num_vars=None, length=None, tagged=False):
    """coeffs: list of (a, b) canonical ints.  Returns the uncompressed wire image."""
    n = len(coeffs)
    lo, hi = chunk_range if chunk_range is not None else (0, n)
    if length is None:
        length = hi - lo
    if num_vars is None:
        num_vars = max(length, 1).bit_length() - 1
    out = b"\x01" if tagged else b""
    out += struct.pack("<Q", num_vars) + _shares(coeffs) + _shares(bound_coeffs)
    out += b"\x00" if scratch is None else b"\x01" + _shares(scratch)
    out += struct.pack("<QQQ", length, lo, hi)
    return out


def deserialize_rep3_dense(raw, tagged=False):
    """Inverse of serialize_rep3_dense: (coeffs within chunk_range, bytes consumed).  Raises ValueError on a bad image."""
    off = 0
    if tagged:
        if raw[0] != 1:
            raise ValueError("not the Shared variant")
        off = 1

    def u64():
        nonlocal off
        if off + 8 > len(raw):
            raise ValueError("unexpected end of input")
        (v,) = struct.unpack_from("<Q", raw, off)
        off += 8
        return v

    def shares():
        nonlocal off
        cnt = u64()
        if off + 64 * cnt > len(raw):
            raise ValueError("unexpected end of input")
        vec = []
        for _ in range(cnt):
            a = int.from_bytes(raw[off:off + 32], "little")
            b = int.from_bytes(raw[off + 32:off + 64], "little")
            if a >= R or b >= R:
                raise ValueError("field element not below the modulus")
            vec.append((a, b))
            off += 64
        return vec

    u64()
    coeffs = shares()
    shares()
    if off >= len(raw):
        raise ValueError("unexpected end of input")
    tag = raw[off]
    off += 1
    if tag > 1:
        raise ValueError("bad Option tag")
    if tag:
        shares()
    u64()
    lo, hi = u64(), u64()
    if lo > hi or hi > len(coeffs):
        raise ValueError("chunk_range outside coeffs")
    return coeffs[lo:hi], off


# ---------------------------------------------------------------- linear combination / evaluation

def linear_combination(polys, coeffs, party):
    """polys: list of ("shared", [(a, b), ...]) or ("public", [v, ...]) with canonical ints (small-scalar polynomials
    are public polynomials whose values happen to be small; negative i64 values are passed mod r).
    Returns ("shared", [(a, b)]) or ("public", [v])."""
    n = max(len(p[1]) for p in polys)
    any_shared = any(p[0] == "shared" for p in polys)
    out = []
    for i in range(n):
        sa = sb = pub = 0
        covered = False
        for (kind, vals), c in zip(polys, coeffs):
            if i >= len(vals):
                continue
            if kind == "shared":
                covered = True
                sa += vals[i][0] * c
                sb += vals[i][1] * c
            else:
                pub += vals[i] * c
        if not any_shared:
            out.append(pub % R)
            continue
        if not covered:
            raise ValueError("Not an arithmetic share")  # as_shared() on SharedOrPublic::Public
        if party == 0:
            sa += pub
        elif party == 1:
            sb += pub
        out.append((sa % R, sb % R))
    return ("shared" if any_shared else "public", out)


def evaluate_at_chi(poly, chis):
    kind, vals = poly
    if len(vals) != len(chis):
        raise ValueError("zip_eq: lengths differ")
    if kind == "shared":
        return sum((a + b) * TWO_INV % R * chi for (a, b), chi in zip(vals, chis)) % R
    return sum(v * chi for v, chi in zip(vals, chis)) % R


def pair_sums(points):
    """points: affine tuples or None; S[b] = P[2b] + P[2b+1]."""
    return [pyref.add(points[2 * b], points[2 * b + 1]) for b in range(len(points) // 2)]


def open_quotients(evals, point):
    """The scalar side of open(): per level the quotient vector q (NOT duplicated) and the final evaluation."""
    r = [e % R for e in evals]
    qs = []
    for t in point:
        half = len(r) // 2
        qs.append([(r[2 * b + 1] - r[2 * b]) % R for b in range(half)])
        r = [(r[2 * b] * (1 - t) + r[2 * b + 1] * t) % R for b in range(half)]
    return qs, r[0]


# ---------------------------------------------------------------- eq tables, co-spartan's batched opening worker

def eq_evals(point, msb_first=False):
    """chi[b] = prod_i (bit_i(b) ? t_i : 1 - t_i) with t_i = point[i] (lsb first) or point[nv - 1 - i] (msb first)."""
    nv = len(point)
    ts = [point[nv - 1 - i] if msb_first else point[i] for i in range(nv)]
    chi = [1]
    for t in ts:  # variable i doubles the table: index bit i is the new top bit
        chi = [c * (1 - t) % R for c in chi] + [c * t % R for c in chi]
    return chi


def dmle_evaluate(evals, point):
    """DenseMultilinearExtension::evaluate(point): fix variable i (index bit i) to point[i], as worker.rs:793-798 folds."""
    r = [e % R for e in evals]
    if len(r) != 1 << len(point):
        raise ValueError("invalid size of partial point")
    for t in point:
        r = [(r[2 * b] * (1 - t) + r[2 * b + 1] * t) % R for b in range(len(r) // 2)]
    return r[0]


def aggregate_poly(eta, polys):
    """utils.rs:85-107: evals[i] += x * p[i] for the polynomials in turn, x *= eta; zip stops at the shorter vector."""
    n = 1 << max(max(len(p), 1).bit_length() - 1 for p in polys)
    out = [0] * n
    x = 1
    for p in polys:
        for i, v in enumerate(p[:n]):
            out[i] = (out[i] + x * v) % R
        x = x * eta % R
    return out


def distributed_batch_open_poly_worker(polys, point, eta, num_comms):
    """worker.rs:745-772 without the group side: (quotient vectors per level - NOT duplicated -, val, evals)."""
    agg = aggregate_poly(eta, polys[:num_comms])
    qs, val = open_quotients(agg, point)
    return qs, val, [dmle_evaluate(p, point) for p in polys]
