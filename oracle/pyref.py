"""Python big-integer restatement of BN254 G1 arithmetic and of a multi-scalar multiplication.

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ may be imported by the product path
(co-zkvms_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it, and only as the checker or the timed CPU baseline.

PARITY UNPINNED: the reference (ChainSafe/co-zkvms) holds no golden vector for this path.
Its MSM lives in un-vendored crates (jolt-core 0.1.0 @ nulltea/jolt cd50b476,
ark-ec/ark-ff 0.5.0 @ a16z/arkworks-algebra 4ae5018, ark-bn254 0.5.0) that cannot be built
here (no Rust toolchain). What this file restates is the *mathematical* contract of the
reference call sites:
  co-jolt/src/poly/commitment/pst13.rs:282-296  commit       = sum_i evals[i] * powers_of_g[0][i], .into_affine()
  co-jolt/src/poly/commitment/pst13.rs:299-331  batch_commit = the same for k polynomials, one base prefix
  co-jolt/src/poly/commitment/pst13.rs:428-474  open         = nv MSMs over q[k][x>>1] (duplicated scalars)
  co-noir-spartan/co-spartan/src/worker.rs:577-590, 774-809  MultilinearPC::commit / msm_bigint
Every caller normalises with .into_affine(), so the affine (x, y, infinity) triple is the
canonical, algorithm-independent output that parity is checked on.

The arithmetic here is deliberately the slowest and most obviously correct form: affine
chord-and-tangent with a modular inverse per step, on Python ints. It is the third independent
implementation next to oracle/bn254.c (4x64 Montgomery, Jacobian) and the CUDA engine
(8x32 Montgomery, XYZZ), and is what the golden fixtures under tests/golden/ are made with
(tests/golden/make_golden.py).
"""

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # base field Fq
R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # scalar field Fr
B_COEFF = 3
G1 = (1, 2)
MASK64 = (1 << 64) - 1
MONT_R = 1 << 256

# ---------------------------------------------------------------- field / curve (affine, None = infinity)


def is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def mul(k, pt):
    k %= R_ORDER
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


def msm_naive(scalars, bases):
    """sum_i scalars[i] * bases[i]; truncates to the shorter slice like ark_ec::msm_bigint."""
    acc = None
    for s, b in zip(scalars, bases):
        acc = add(acc, mul(s, b))
    return acc


def msm_windowed(scalars, bases, c=8):
    """A plain (unsigned-digit) bucket method on Python ints: faster than msm_naive for n ~ 2^10."""
    nwin = (254 + c - 1) // c
    total = None
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = add(total, total)
        buckets = [None] * (1 << c)
        for s, b in zip(scalars, bases):
            d = ((s % R_ORDER) >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = add(buckets[d], b)
        run = None
        acc = None
        for d in range((1 << c) - 1, 0, -1):
            run = add(run, buckets[d])
            acc = add(acc, run)
        total = add(total, acc)
    return total


# ---------------------------------------------------------------- deterministic synthetic inputs
# counter-based SplitMix64, identical in Python, C (oracle/bn254.c) and CUDA
# (co-zkvms_b200/csrc/testgen.cu).


def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def limb(seed, i, j):
    return mix64((mix64(seed) + 4 * i + j) & MASK64)


def raw254(seed, i):
    v = 0
    for j in range(4):
        v |= limb(seed, i, j) << (64 * j)
    return v & ((1 << 254) - 1)


def scalar_uniform(seed, i):
    v = raw254(seed, i)
    return v - R_ORDER if v >= R_ORDER else v


def sqrt_fq(a):
    """p = 3 mod 4: candidate root a^((p+1)/4); returns None for non-residues."""
    y = pow(a, (P + 1) // 4, P)
    return y if y * y % P == a % P else None


def base_point(seed, i):
    """Try-and-increment: first x >= H(seed, i) with x^3+3 a square; even canonical y."""
    x = raw254(seed, i)
    if x >= P:
        x -= P
    while True:
        y = sqrt_fq((x * x * x + B_COEFF) % P)
        if y is not None:
            if y & 1:
                y = P - y
            return (x, y)
        x = (x + 1) % P


def scalars(dist, seed, n):
    """Scalar distributions of SURVEY.md section 8(d):
    uniform   - co-spartan shares (mpc-core/src/protocols/rep3/poly.rs:94-95)
    const     - co-jolt party 0/1 share a (dense_mlpoly.rs:567-585: the mask is ONE random value repeated)
    wminus    - co-jolt party 2: (w_i - c0 - c1) mod r with w_i < 2^32
    dup       - PST13 open(): every quotient value appears twice (pst13.rs:459)
    small16   - public polynomials with u16 coefficients
    zero_half - padded traces (dense_mlpoly.rs:59-64): second half all zero
    """
    if dist == "uniform":
        return [scalar_uniform(seed, i) for i in range(n)]
    if dist == "const":
        c0 = scalar_uniform(seed, 0)
        return [c0] * n
    if dist == "wminus":
        c0 = scalar_uniform(seed, 0)
        c1 = scalar_uniform(seed, 1)
        return [((limb(seed, i + 2, 0) & 0xFFFFFFFF) - c0 - c1) % R_ORDER for i in range(n)]
    if dist == "dup":
        return [scalar_uniform(seed, i >> 1) for i in range(n)]
    if dist == "small16":
        return [limb(seed, i, 0) & 0xFFFF for i in range(n)]
    if dist == "zero_half":
        return [scalar_uniform(seed, i) if i < (n + 1) // 2 else 0 for i in range(n)]
    raise ValueError(dist)


DISTS = ("uniform", "const", "wminus", "dup", "small16", "zero_half")


# ---------------------------------------------------------------- Montgomery helpers (wire format of the C ABI)


def to_mont(x, mod):
    return x * MONT_R % mod


def from_mont(x, mod):
    return x * pow(MONT_R, -1, mod) % mod


def limbs_le(x, n=4):
    return [(x >> (64 * k)) & MASK64 for k in range(n)]
