"""BN254 optimal ate pairing on Python integers - TEST INFRASTRUCTURE ONLY (see oracle/bn254.h).

Why it exists: the reference's acceptance test for an opening is the pairing check
(co-jolt/src/poly/commitment/pst13.rs:536-545: PST13::prove -> PST13::verify -> MultilinearPC::check; end to end in
co-jolt/examples/rep3_jolt.rs:311).  `verify_opening` below evaluates that verifier equation, so proofs produced by the
engine (cozk_pst13_open / cozk_pst13_open_poly) are pinned by something other than a restatement of open() itself: a
wrong point order, level order or base pairing in the prover would not satisfy it.

MultilinearPC::check lives in an un-vendored crate (ark-poly-commit 0.5.0 @ nulltea/poly-commit 40eb68d); what is restated
here is the published PST13 equation in the form the reference's data types fix (proof elements in G1 - open() multiplies
powers_of_g[i], pst13.rs:461-467 - and the verifier key's h, h_mask in G2):

    e(C - value * g, h)  =  prod_i  e(proof_i, h_mask_i - point_i * h),        h_mask_i = t_i * h

Field tower as in EIP-197 implementations: Fp2 = Fp[i]/(i^2 + 1), Fp12 = Fp[w]/(w^12 - 18 w^6 + 82) with w^6 = 9 + i;
G2 is the D-type sextic twist y^2 = x^3 + 3/(9 + i) over Fp2.  Everything is plain big-integer arithmetic.
"""
from . import pyref

P = pyref.P
R = pyref.R_ORDER
ATE_LOOP = 29793968203157093288  # 6u + 2, u = 4965661367192848881

# ------------------------------------------------------------------------------------------------ Fp2 (pairs of ints)


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return (-a[0] % P, -a[1] % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_muli(a, k):
    return (a[0] * k % P, a[1] * k % P)


def f2_conj(a):
    return (a[0], -a[1] % P)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], P - 2, P)
    return (a[0] * d % P, -a[1] * d % P)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


XI = (9, 1)
B2 = f2_muli(f2_inv(XI), 3)  # twist curve constant 3 / (9 + i)
G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
       11559732032986387107991004021392285783925812861821192530917403151452391805634),
      (8495653923123431417604973247489272438418190587263600148770280649306958101930,
       4082367875863433681332203403145435568316851327593401208105741076214120093531))
# Frobenius on twist points: (x, y) -> (conj(x) * GAMMA12, conj(y) * GAMMA13); squared: (x * GAMMA22, y * GAMMA23)
GAMMA12 = f2_pow(XI, (P - 1) // 3)
GAMMA13 = f2_pow(XI, (P - 1) // 2)
GAMMA22 = f2_pow(XI, (P * P - 1) // 3)
GAMMA23 = f2_pow(XI, (P * P - 1) // 2)

# ------------------------------------------------------------------------------------------------ G2 (affine over Fp2, None = infinity)


def g2_is_on_curve(q):
    if q is None:
        return True
    x, y = q
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), B2)


def g2_neg(q):
    return None if q is None else (q[0], f2_neg(q[1]))


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if a[1] != b[1] or a[1] == (0, 0):
            return None
        lam = f2_mul(f2_muli(f2_mul(a[0], a[0]), 3), f2_inv(f2_muli(a[1], 2)))
    else:
        lam = f2_mul(f2_sub(b[1], a[1]), f2_inv(f2_sub(b[0], a[0])))
    x = f2_sub(f2_sub(f2_mul(lam, lam), a[0]), b[0])
    return (x, f2_sub(f2_mul(lam, f2_sub(a[0], x)), a[1]))


def g2_mul(q, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, q)
        q = g2_add(q, q)
        k >>= 1
    return acc


# ------------------------------------------------------------------------------------------------ Fp12 (12 coefficients of w)
F12_ONE = [1] + [0] * 11


def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):  # w^12 = 18 w^6 - 82
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


def _line(T, Q, Pt):
    """Line through the untwisted images of T and Q (tangent when T == Q) evaluated at the G1 point Pt, as an Fp12
    element; returns (line, T + Q).  With psi(x, y) = (x w^2, y w^3):  l(P) = yP - lam xP w + (lam xT - yT) w^3, and an
    Fp2 value a + b i sits at (a - 9 b) w^k + b w^(k+6) because i = w^6 - 9."""
    xT, yT = T
    if T == Q:
        lam = f2_mul(f2_muli(f2_mul(xT, xT), 3), f2_inv(f2_muli(yT, 2)))
    else:
        lam = f2_mul(f2_sub(Q[1], yT), f2_inv(f2_sub(Q[0], xT)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), xT), Q[0])
    y3 = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT)
    c1 = f2_muli(lam, -Pt[0] % P)            # coefficient of w
    c3 = f2_sub(f2_mul(lam, xT), yT)         # coefficient of w^3
    l = [0] * 12
    l[0] = Pt[1] % P
    l[1] = (c1[0] - 9 * c1[1]) % P
    l[7] = c1[1]
    l[3] = (c3[0] - 9 * c3[1]) % P
    l[9] = c3[1]
    return l, (x3, y3)


def miller_loop(Pt, Q):
    """f_{6u+2,Q}(P) times the two Frobenius lines; Pt in G1 (affine ints), Q in G2 (affine Fp2).  No final exponentiation."""
    if Pt is None or Q is None:
        return F12_ONE
    f = F12_ONE
    T = Q
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        l, T2 = _line(T, T, Pt)
        f = f12_mul(f12_mul(f, f), l)
        T = T2
        if (ATE_LOOP >> i) & 1:
            l, T = _line(T, Q, Pt)
            f = f12_mul(f, l)
    Q1 = (f2_mul(f2_conj(Q[0]), GAMMA12), f2_mul(f2_conj(Q[1]), GAMMA13))
    nQ2 = (f2_mul(Q[0], GAMMA22), f2_neg(f2_mul(Q[1], GAMMA23)))
    l, T = _line(T, Q1, Pt)
    f = f12_mul(f, l)
    l, T = _line(T, nQ2, Pt)
    return f12_mul(f, l)


FINAL_EXP = (P ** 12 - 1) // R


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def pairing(Pt, Q):
    return final_exponentiation(miller_loop(Pt, Q))


def pairing_product_is_one(pairs):
    """prod e(P_k, Q_k) == 1, with one shared final exponentiation."""
    f = F12_ONE
    for Pt, Q in pairs:
        f = f12_mul(f, miller_loop(Pt, Q))
    return final_exponentiation(f) == F12_ONE


# ------------------------------------------------------------------------------------------------ PST13 / MultilinearPC
def setup(nv, seed):
    """An eq-basis SRS as MultilinearPC::setup builds it (co-jolt/src/poly/commitment/pst13.rs:49-62, :233-250), from a
    seeded trapdoor t_0 .. t_{nv-1}:  powers_of_g[i][x] = eq(x; t_i, .., t_{nv-1}) * g  (bit j of x pairs with t_{i+j}: the
    pairing of level i with point[i] and of the LOWEST index bit that open() folds first, pst13.rs:451-459), i = 0 .. nv-1;
    h = G2 generator, h_mask[i] = t_i * h.  Returns (levels: list of lists of affine G1 points, vk dict, trapdoor)."""
    t = [pyref.scalar_uniform(seed, 1000 + i) or 1 for i in range(nv)]
    g = pyref.G1
    levels = []
    for i in range(nv):
        k = nv - i
        pts = []
        for x in range(1 << k):
            e = 1
            for j in range(k):
                e = e * (t[i + j] if (x >> j) & 1 else (1 - t[i + j])) % R
            pts.append(pyref.mul(e, g))
        levels.append(pts)
    vk = {"nv": nv, "g": g, "h": G2, "h_mask": [g2_mul(G2, ti) for ti in t]}
    return levels, vk, t


def verify_opening(vk, commitment, point, value, proofs):
    """MultilinearPC::check: e(C - value g, h) == prod_i e(proof_i, h_mask_i - point_i h).  commitment / proofs: affine G1
    points (None = infinity); point: the nv field elements in the order open() consumed them; value: the claimed f(point)."""
    nv = vk["nv"]
    assert len(point) == nv and len(proofs) == nv
    left = pyref.add(commitment, pyref.neg(pyref.mul(value % R, vk["g"])))
    pairs = [(pyref.neg(left), vk["h"])]
    for i in range(nv):
        rhs = g2_add(vk["h_mask"][i], g2_neg(g2_mul(vk["h"], point[i] % R)))
        pairs.append((proofs[i], rhs))
    return pairing_product_is_one(pairs)
