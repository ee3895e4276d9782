// The hot mixed addition of the bucket-accumulation kernel on the 9 x 29-bit field (field29.cuh), written with lazy
// additive operations and proven value bounds instead of a modular reduction after every add/sub.
//
// Accumulator invariant between two calls (all limbs normalised):
//     X < 8p,  Y < 4p,  ZZ < 2p,  ZZZ < 2p;   identity <=> all limbs of ZZ are zero.
// Operand: affine point with x < 2p, y < 2p (table entries are < p; a negated y is p - y <= p).
//
//   U2 = x*ZZ, S2 = y*ZZZ                      < 2p          (mul returns < 2p for operands < 2^257)
//   P  = U2 - X + 8p                           < 10p < 2^257
//   R  = S2 - Y + 4p                           < 6p
//   PP = P^2, PPP = P*PP, Q = X*PP             < 2p
//   X3 = R^2 - PPP - 2Q + 6p                   < 8p          (PPP + 2Q < 6p keeps it positive)
//   T  = Q - X3 + 8p                           < 10p
//   Y3 = R*T - Y*PPP + 2p                      < 4p
//   ZZ3 = ZZ*PP, ZZZ3 = ZZZ*PPP                < 2p
// Every product has both operands < 10p, so a*b < 100 p^2 < 2^261 p as mul() requires.
// P = 0 mod p (the operand equals +-acc) is detected on PP, which is < 2p: PP in {0, p}.
#pragma once
#include "field29.cuh"

namespace cozk {
namespace f29 {

struct affine29 {
    fe x, y;
};
struct xyzz29 {
    fe X, Y, ZZ, ZZZ;
};

COZK_HD xyzz29 identity29() {
    xyzz29 r;
    r.X = zero();
    r.Y = zero();
    r.ZZ = zero();
    r.ZZZ = zero();
    return r;
}
COZK_HD bool is_identity(const xyzz29& p) { return is_zero_limbs(p.ZZ); }

// bring a lazy accumulator back to the storage invariant: every coordinate in [0, 2p)
COZK_HD xyzz29 normalize29(const xyzz29& a) {
    xyzz29 r;
    r.X = csub<2>(csub<4>(a.X));
    r.Y = csub<2>(a.Y);
    r.ZZ = a.ZZ;
    r.ZZZ = a.ZZZ;
    return r;
}

// 2 * (x, y), x, y < 2p  ->  all coordinates < 2p      (mdbl-2008-s-1, a = 0)
COZK_HD xyzz29 dbl_affine29(const affine29& p) {
    xyzz29 r;
    fe U = dbl_mod(p.y);
    fe V = sqr(U);
    fe W = mul(U, V);
    fe S = mul(p.x, V);
    fe xx = sqr(p.x);
    fe M = add_mod(dbl_mod(xx), xx);
    r.X = sub_mod(sqr(M), dbl_mod(S));
    r.Y = sub_mod(mul(M, sub_mod(S, r.X)), mul(W, p.y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}

// acc + q with the lazy bounds of the header comment (madd-2008-s)
COZK_HD xyzz29 madd_lazy(const xyzz29& a, const affine29& q) {
    if (is_identity(a)) {
        xyzz29 r;
        r.X = q.x;
        r.Y = q.y;
        r.ZZ = one();
        r.ZZZ = one();
        return r;
    }
    fe U2 = mul(q.x, a.ZZ);
    fe S2 = mul(q.y, a.ZZZ);
    fe P = sub<8>(U2, a.X);
    fe R = sub<4>(S2, a.Y);
    fe PP = sqr(P);
    if (is_zero_mod_p(PP)) {
        // q = +-acc: R = 0 mod p means the same point (double it), otherwise the sum is the identity
        if (is_zero_mod_p(csub<2>(csub<4>(R)))) return dbl_affine29(q);
        return identity29();
    }
    xyzz29 r;
    fe PPP = mul(P, PP);
    fe Q = mul(a.X, PP);
    r.X = sub2<6>(sqr(R), PPP, add_nocarry(Q, Q));
    fe T = sub<8>(Q, r.X);
    r.Y = sub<2>(mul(R, T), mul(a.Y, PPP));
    r.ZZ = mul(a.ZZ, PP);
    r.ZZZ = mul(a.ZZZ, PPP);
    return r;
}

// wire point (72 B, arkworks-form affine)  <->  xyzz29
COZK_HD xyzz29 from_wire29(const uint8_t* in72) {
    if (in72[64]) return identity29();
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in72);
    uint32_t wx[8], wy[8];
    for (int i = 0; i < 8; ++i) {
        wx[i] = w[i];
        wy[i] = w[8 + i];
    }
    xyzz29 r;
    r.X = from_ark(wx);
    r.Y = from_ark(wy);
    r.ZZ = one();
    r.ZZZ = one();
    return r;
}
// coordinates must be < 2p (normalize29 first for a lazy accumulator)
COZK_HD void to_wire29(const xyzz29& p, uint8_t* out72) {
    uint32_t* w = reinterpret_cast<uint32_t*>(out72);
    for (int i = 0; i < 18; ++i) w[i] = 0;
    if (is_identity(p)) {
        out72[64] = 1;
        return;
    }
    fe I = inv(mul(p.ZZ, p.ZZZ));
    fe x = mul(p.X, mul(I, p.ZZZ));
    fe y = mul(p.Y, mul(I, p.ZZ));
    uint32_t wx[8], wy[8];
    to_ark(x, wx);
    to_ark(y, wy);
    for (int i = 0; i < 8; ++i) {
        w[i] = wx[i];
        w[8 + i] = wy[i];
    }
}

}  // namespace f29
}  // namespace cozk
