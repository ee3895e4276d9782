// BN254 G1 (y^2 = x^3 + 3) group law for the bucket method.
//
//   affine  - a finite SRS point as the reference stores it (ark_ec short_weierstrass::Affine{x, y}; the `infinity`
//             flag travels separately, see include/cozk_msm.h), 64 B.
//   xyzz    - extended Jacobian accumulator (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ = 0 is the
//             identity.  128 B.  Mixed addition costs 8M + 2S, full addition 12M + 2S, doubling 6M + 3S; in each of
//             them Y3 = u*v - w*z is ONE fused product pair (fq_mul2: two products, one Montgomery reduction).
//
// The reference's CPU path accumulates in Jacobian `Projective` and normalises with `.into_affine()`
// (pst13.rs:294, :328, :469); coordinates differ, the group element - and so the affine output - does not.
// All exceptional cases are handled exactly (P + P, P + (-P), identity operands): degenerate co-jolt shares and
// duplicated SRS entries make them reachable.
#pragma once
#include "field.cuh"

namespace cozk {

struct alignas(16) affine {
    fq x, y;
};
struct alignas(16) xyzz {
    fq X, Y, ZZ, ZZZ;
};

COZK_HD xyzz xyzz_identity() {
    xyzz r;
    r.X = fq_zero();
    r.Y = fq_zero();
    r.ZZ = fq_zero();
    r.ZZZ = fq_zero();
    return r;
}
COZK_HD bool xyzz_is_identity(const xyzz& p) { return fq_is_zero(p.ZZ); }
COZK_HD xyzz xyzz_from_affine(const affine& p) {
    xyzz r;
    r.X = p.x;
    r.Y = p.y;
    r.ZZ = fq_one();
    r.ZZZ = fq_one();
    return r;
}
COZK_HD xyzz xyzz_neg(const xyzz& p) {
    xyzz r = p;
    r.Y = fq_neg(p.Y);
    return r;
}

// 2 * (x, y) -> XYZZ   (mdbl-2008-s-1, a = 0)
COZK_HD xyzz xyzz_dbl_affine(const affine& p) {
    xyzz r;
    fq U = fq_dbl(p.y);
    fq V = fq_sqr(U);
    fq W = fq_mul(U, V);
    fq S = fq_mul(p.x, V);
    fq xx = fq_sqr(p.x);
    fq M = fq_add(fq_dbl(xx), xx);
    r.X = fq_sub(fq_sqr(M), fq_dbl(S));
    r.Y = fq_mul2(M, fq_sub(S, r.X), W, fq_neg(p.y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}

// 2 * P   (dbl-2008-s-1, a = 0).  Y = 0 (a 2-torsion point; none on BN254 G1) gives ZZ = 0 = identity by itself.
COZK_HD xyzz xyzz_dbl(const xyzz& p) {
    xyzz r;
    fq U = fq_dbl(p.Y);
    fq V = fq_sqr(U);
    fq W = fq_mul(U, V);
    fq S = fq_mul(p.X, V);
    fq xx = fq_sqr(p.X);
    fq M = fq_add(fq_dbl(xx), xx);
    r.X = fq_sub(fq_sqr(M), fq_dbl(S));
    r.Y = fq_mul2(M, fq_sub(S, r.X), W, fq_neg(p.Y));
    r.ZZ = fq_mul(V, p.ZZ);
    r.ZZZ = fq_mul(W, p.ZZZ);
    return r;
}

// acc + (x, y)   (madd-2008-s)
COZK_HD xyzz xyzz_madd(const xyzz& a, const affine& q) {
    if (xyzz_is_identity(a)) return xyzz_from_affine(q);
    fq U2 = fq_mul(q.x, a.ZZ);
    fq S2 = fq_mul(q.y, a.ZZZ);
    fq P = fq_sub(U2, a.X);
    fq R = fq_sub(S2, a.Y);
    if (fq_is_zero(P)) {
        if (fq_is_zero(R)) return xyzz_dbl_affine(q);
        return xyzz_identity();
    }
    xyzz r;
    fq PP = fq_sqr(P);
    fq PPP = fq_mul(P, PP);
    fq Q = fq_mul(a.X, PP);
    r.X = fq_sub(fq_sub(fq_sqr(R), PPP), fq_dbl(Q));
    r.Y = fq_mul2(R, fq_sub(Q, r.X), fq_neg(a.Y), PPP);
    r.ZZ = fq_mul(a.ZZ, PP);
    r.ZZZ = fq_mul(a.ZZZ, PPP);
    return r;
}

// a + b   (add-2008-s)
COZK_HD xyzz xyzz_add(const xyzz& a, const xyzz& b) {
    if (xyzz_is_identity(a)) return b;
    if (xyzz_is_identity(b)) return a;
    fq U1 = fq_mul(a.X, b.ZZ);
    fq U2 = fq_mul(b.X, a.ZZ);
    fq S1 = fq_mul(a.Y, b.ZZZ);
    fq S2 = fq_mul(b.Y, a.ZZZ);
    fq P = fq_sub(U2, U1);
    fq R = fq_sub(S2, S1);
    if (fq_is_zero(P)) {
        if (fq_is_zero(R)) return xyzz_dbl(a);
        return xyzz_identity();
    }
    xyzz r;
    fq PP = fq_sqr(P);
    fq PPP = fq_mul(P, PP);
    fq Q = fq_mul(U1, PP);
    r.X = fq_sub(fq_sub(fq_sqr(R), PPP), fq_dbl(Q));
    r.Y = fq_mul2(R, fq_sub(Q, r.X), fq_neg(S1), PPP);
    r.ZZ = fq_mul(fq_mul(a.ZZ, b.ZZ), PP);
    r.ZZZ = fq_mul(fq_mul(a.ZZZ, b.ZZZ), PPP);
    return r;
}

// Normalise to the C ABI's 72-byte result: x[32] || y[32] (Fq Montgomery, canonical) || infinity || pad[7].
// One inversion: I = 1/(ZZ*ZZZ); 1/ZZ = I*ZZZ; 1/ZZZ = I*ZZ.
COZK_HD void xyzz_to_wire(const xyzz& p, uint8_t* out72) {
    uint32_t* w = reinterpret_cast<uint32_t*>(out72);
    if (xyzz_is_identity(p)) {
        for (int i = 0; i < 18; ++i) w[i] = 0;
        out72[64] = 1;
        return;
    }
    fq I = fq_inv(fq_mul(p.ZZ, p.ZZZ));
    fq x = fq_mul(p.X, fq_mul(I, p.ZZZ));
    fq y = fq_mul(p.Y, fq_mul(I, p.ZZ));
    for (int i = 0; i < 8; ++i) {
        w[i] = x.v[i];
        w[8 + i] = y.v[i];
    }
    w[16] = 0;
    w[17] = 0;
}

// 72-byte wire point -> XYZZ (used by the host-side combine of per-GPU partial sums)
COZK_HD xyzz xyzz_from_wire(const uint8_t* in72) {
    if (in72[64]) return xyzz_identity();
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in72);
    affine a;
    for (int i = 0; i < 8; ++i) {
        a.x.v[i] = w[i];
        a.y.v[i] = w[8 + i];
    }
    return xyzz_from_affine(a);
}

}  // namespace cozk
