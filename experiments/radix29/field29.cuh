// BN254 Fq on 9 limbs of 29 bits (in 32-bit registers), Montgomery radix R' = 2^261: the representation the hot
// kernels compute in.
//
// Why not 8 x 32: measured on B200 (tools/microbench.py, profiles/), IMAD.WIDE.U32 issues at the full integer-multiply
// rate (64 lanes/clk/SM) only WITHOUT a carry predicate; the carry-chained form (IMAD.WIDE.U32.X, what mad.lo.cc /
// madc.hi.cc pairs compile to) runs at half that rate, and a 32-bit-limb Montgomery product is nothing but such
// chains (136 of them = 544 pipe cycles per warp).  With 29-bit limbs a column of a product-scanning Montgomery
// multiplication - up to 9 a_i*b_j plus 9 m_i*p_j terms of < 2^58 each - fits a 64-bit accumulator with no carry at
// all, so the whole product is 81 + 81 plain IMAD.WIDE + 9 IMAD = 171 multiply-adds = 342 pipe cycles, and the
// carry handling (one 64-bit shift and a mask per column) moves to the otherwise idle ALU pipe.
//
// Value discipline.  R' = 2^261 > 64 p, so mul() accepts any operands whose product is < 2^261 * p (in particular
// both < 2^257, ~ 10 p) and returns a value < 2p with normalised limbs.  Additive operations come in two flavours:
//   *_mod   keep the invariant "value in [0, 2p), limbs normalised" (used by the generic group law);
//   lazy    add / sub<K> only propagate carries (sub adds K*p to stay positive); the caller tracks the bound.  The
//           mixed addition of the bucket-accumulation kernel is written with these and proven bounds (curve.cuh).
// Limb i < 8 is normalised when < 2^29; limb 8 holds the rest (value < 2^261 <=> limb 8 < 2^29).
#pragma once
#include <cstdint>

#include "../../co-zkvms_b200/csrc/field.cuh"
#include "field29_consts.inc"

namespace cozk {
namespace f29 {

constexpr int NL = 9;
constexpr uint32_t LB = 29;
constexpr uint32_t MASK = (1u << LB) - 1u;

struct fe {
    uint32_t v[NL];
};

// 32 x 32 + 64 -> 64 multiply-add: exactly one IMAD.WIDE.U32 (no carry predicate).  Spelled in PTX on the device
// because (uint64_t)a * CONSTANT + c makes nvcc emit a 64-bit constant multiply with a dead high-word add.
COZK_HD uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#else
    return (uint64_t)a * b + c;
#endif
}

COZK_HD fe zero() {
    fe r;
#pragma unroll
    for (int i = 0; i < NL; ++i) r.v[i] = 0;
    return r;
}
COZK_HD fe one() {
    const uint32_t c[NL] = F29_ONE;
    fe r;
#pragma unroll
    for (int i = 0; i < NL; ++i) r.v[i] = c[i];
    return r;
}
COZK_HD bool is_zero_limbs(const fe& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i) o |= a.v[i];
    return o == 0;
}
// value in [0, 2p) with normalised limbs: is it 0 mod p?
COZK_HD bool is_zero_mod_p(const fe& a) {
    const uint32_t P[NL] = F29_P;
    uint32_t o = 0, q = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        o |= a.v[i];
        q |= a.v[i] ^ P[i];
    }
    return o == 0 || q == 0;
}

// Montgomery product a*b/2^261 mod p, product scanning.  Requires: limbs of b normalised, limbs of a < 2^30
// (so a may be an un-normalised sum of two normalised values), a*b < 2^261 * p.  Returns value < 2p, normalised.
// Column bound: 9 * 2^59 + 9 * 2^58 + carry < 2^63.
COZK_HD fe mul(const fe& a, const fe& b) {
    const uint32_t P[NL] = F29_P;
    uint32_t m[NL];
    fe r;
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i <= k; ++i) s = madw(a.v[i], b.v[k - i], s);
#pragma unroll
        for (int i = 0; i < k; ++i) acc = madw(m[i], P[k - i], acc);
        acc += s;
        m[k] = ((uint32_t)acc * F29_N0) & MASK;
        acc = madw(m[k], P[0], acc);
        acc >>= LB;
    }
#pragma unroll
    for (int k = NL; k < 2 * NL - 1; ++k) {
        uint64_t s = 0;
#pragma unroll
        for (int i = k - NL + 1; i < NL; ++i) s = madw(a.v[i], b.v[k - i], s);
#pragma unroll
        for (int i = k - NL + 1; i < NL; ++i) acc = madw(m[i], P[k - i], acc);
        acc += s;
        r.v[k - NL] = (uint32_t)acc & MASK;
        acc >>= LB;
    }
    r.v[NL - 1] = (uint32_t)acc;
    return r;
}

// a^2: the cross products are computed once and doubled: 45 + 81 + 9 multiply-adds.  Limbs of a < 2^30 allowed
// when the value bound of mul() holds (2 * 2^60 * 4 + ... stays < 2^64 only for normalised limbs, so: normalised).
COZK_HD fe sqr(const fe& a) {
    const uint32_t P[NL] = F29_P;
    uint32_t m[NL];
    fe r;
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < 2 * NL - 1; ++k) {
        uint64_t s = 0;
        const int lo = k < NL ? 0 : k - NL + 1;
        const int hi = k < NL ? k : NL - 1;
#pragma unroll
        for (int i = lo; i <= hi; ++i) {
            int j = k - i;
            if (i < j) s = madw(a.v[i], a.v[j], s);
        }
        s <<= 1;
        if ((k & 1) == 0) s = madw(a.v[k / 2], a.v[k / 2], s);
        if (k < NL) {
#pragma unroll
            for (int i = 0; i < k; ++i) acc = madw(m[i], P[k - i], acc);
            acc += s;
            m[k] = ((uint32_t)acc * F29_N0) & MASK;
            acc = madw(m[k], P[0], acc);
        } else {
#pragma unroll
            for (int i = k - NL + 1; i < NL; ++i) acc = madw(m[i], P[k - i], acc);
            acc += s;
            r.v[k - NL] = (uint32_t)acc & MASK;
        }
        acc >>= LB;
    }
    r.v[NL - 1] = (uint32_t)acc;
    return r;
}

// carry propagation over signed limb values; the represented value must be >= 0 and < 2^261
COZK_HD fe propagate(const int32_t (&t)[NL]) {
    fe r;
    int32_t c = 0;
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) {
        int32_t x = t[i] + c;
        c = x >> LB;
        r.v[i] = (uint32_t)x & MASK;
    }
    r.v[NL - 1] = (uint32_t)(t[NL - 1] + c);
    return r;
}

// ---- lazy additive operations: carries only, no modular reduction; the caller tracks value bounds
COZK_HD fe add(const fe& a, const fe& b) {
    int32_t t[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) t[i] = (int32_t)(a.v[i] + b.v[i]);
    return propagate(t);
}
// limb-wise sum without carry propagation: limbs < 2^30, usable as the FIRST operand of mul()
COZK_HD fe add_nocarry(const fe& a, const fe& b) {
    fe r;
#pragma unroll
    for (int i = 0; i < NL; ++i) r.v[i] = a.v[i] + b.v[i];
    return r;
}
template <int K>
struct kp;  // limbs of K * p
#define F29_KP(K, NAME)                                     \
    template <>                                             \
    struct kp<K> {                                          \
        COZK_HD static uint32_t limb(int i) {               \
            const uint32_t c[NL] = NAME;                    \
            return c[i];                                    \
        }                                                   \
    };
F29_KP(2, F29_P2)
F29_KP(4, F29_P4)
F29_KP(6, F29_P6)
F29_KP(8, F29_P8)
F29_KP(10, F29_P10)
F29_KP(12, F29_P12)
F29_KP(16, F29_P16)
#undef F29_KP
// a - b + K*p with carries propagated; requires b <= K*p so that the result is >= 0
template <int K>
COZK_HD fe sub(const fe& a, const fe& b) {
    int32_t t[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) t[i] = (int32_t)(a.v[i] + kp<K>::limb(i)) - (int32_t)b.v[i];
    return propagate(t);
}
// a - b - c + K*p
template <int K>
COZK_HD fe sub2(const fe& a, const fe& b, const fe& c) {
    int32_t t[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) t[i] = (int32_t)(a.v[i] + kp<K>::limb(i)) - (int32_t)b.v[i] - (int32_t)c.v[i];
    return propagate(t);
}
// value < 2 * K * p  ->  subtract K*p if that stays >= 0  (one conditional step of a reduction ladder)
template <int K>
COZK_HD fe csub(const fe& a) {
    int32_t t[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) t[i] = (int32_t)a.v[i] - (int32_t)kp<K>::limb(i);
    // signed propagate keeping the sign in the top limb
    uint32_t lo[NL];
    int32_t c = 0;
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) {
        int32_t x = t[i] + c;
        c = x >> LB;
        lo[i] = (uint32_t)x & MASK;
    }
    int32_t top = t[NL - 1] + c;
    bool neg = top < 0;
    fe r;
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) r.v[i] = neg ? a.v[i] : lo[i];
    r.v[NL - 1] = neg ? a.v[NL - 1] : (uint32_t)top;
    return r;
}

// ---- modular flavour: inputs in [0, 2p) normalised, output in [0, 2p) normalised
COZK_HD fe add_mod(const fe& a, const fe& b) { return csub<2>(add(a, b)); }
COZK_HD fe sub_mod(const fe& a, const fe& b) { return csub<2>(sub<2>(a, b)); }
COZK_HD fe dbl_mod(const fe& a) { return add_mod(a, a); }
COZK_HD fe neg_mod(const fe& a) { return csub<2>(sub<2>(zero(), a)); }
// y := cond ? -y : y for y in [0, 2p): 2p - y is in (0, 2p]; csub<2> maps 2p to 0
COZK_HD fe cneg_mod(const fe& a, bool cond) {
    fe n = neg_mod(a), r;
#pragma unroll
    for (int i = 0; i < NL; ++i) r.v[i] = cond ? n.v[i] : a.v[i];
    return r;
}
// [0, 2p) -> [0, p)
COZK_HD fe reduce_full(const fe& a) {
    const uint32_t P[NL] = F29_P;
    int32_t c = 0;
    uint32_t lo[NL];
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) {
        int32_t x = (int32_t)a.v[i] - (int32_t)P[i] + c;
        c = x >> LB;
        lo[i] = (uint32_t)x & MASK;
    }
    int32_t top = (int32_t)a.v[NL - 1] - (int32_t)P[NL - 1] + c;
    bool neg = top < 0;
    fe r;
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) r.v[i] = neg ? a.v[i] : lo[i];
    r.v[NL - 1] = neg ? a.v[NL - 1] : (uint32_t)top;
    return r;
}

// ---- packing: 256-bit little-endian integer in 8 x u32  <->  9 x 29-bit limbs
COZK_HD fe unpack(const uint32_t (&w)[8]) {
    fe r;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int bit = LB * i, q = bit >> 5, off = bit & 31;
        uint64_t x = w[q];
        if (q + 1 < 8) x |= (uint64_t)w[q + 1] << 32;
        r.v[i] = (uint32_t)(x >> off) & MASK;
    }
    return r;
}
// requires normalised limbs and value < 2^256
COZK_HD void pack(const fe& a, uint32_t (&w)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int bit = 32 * j, q = bit / LB, off = bit % LB;
        uint64_t x = (uint64_t)a.v[q] >> off;
        if (q + 1 < NL) x |= (uint64_t)a.v[q + 1] << (LB - off);
        if (q + 2 < NL && 2 * LB - off < 32) x |= (uint64_t)a.v[q + 2] << (2 * LB - off);
        w[j] = (uint32_t)x;
    }
}

// arkworks in-memory Fq (x * 2^256 mod p, canonical, 8 x u32)  ->  internal (x * 2^261 mod p, < 2p)
COZK_HD fe from_ark(const uint32_t (&w)[8]) {
    const uint32_t cin[NL] = F29_C_IN;
    fe c;
#pragma unroll
    for (int i = 0; i < NL; ++i) c.v[i] = cin[i];
    return mul(unpack(w), c);
}
// internal (< 2p)  ->  arkworks in-memory Fq, fully reduced
COZK_HD void to_ark(const fe& a, uint32_t (&w)[8]) {
    const uint32_t cout[NL] = F29_C_OUT;
    fe c;
#pragma unroll
    for (int i = 0; i < NL; ++i) c.v[i] = cout[i];
    pack(reduce_full(mul(a, c)), w);
}

// a^(p-2) for a in [0, 2p); 0 -> 0
COZK_HD fe inv(const fe& a) {
    const uint32_t e[8] = {0xd87cfd45u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    fe acc = one();
    for (int i = 253; i >= 0; --i) {
        acc = sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = mul(acc, a);
    }
    return acc;
}

}  // namespace f29
}  // namespace cozk
