// BN254 base field Fq and scalar field Fr on 8 x 32-bit little-endian limbs, Montgomery form (R = 2^256).
//
// The memory image of an element is identical to arkworks' Fp256<MontBackend<_, 4>> (4 x u64 LE limbs), which is
// what the reference hands to its MSM: bases `&[G1Affine]` and scalars `&[Fr]`
// (co-jolt/src/poly/commitment/pst13.rs:286-292, :319-323).
//
// Device code uses the generated inline-PTX carry chains of field_ptx.inc (tools/gen_field_ptx.py; IMAD.WIDE.U32
// with carry predicates).  Host code (the few group operations the runtime does when it combines per-GPU partial
// sums, and the host-emulation test harness under tests/emul/) uses the portable bodies below.  There is no CPU
// MSM path in this library.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define COZK_HD __host__ __device__ __forceinline__
#define COZK_D __device__ __forceinline__
#else
#define COZK_HD inline
#define COZK_D inline
#endif

namespace cozk {

struct alignas(16) fq {
    uint32_t v[8];
};
typedef fq fr;  // same layout, different modulus; only mul-by-one (from Montgomery) and comparisons are needed

// p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
#define COZK_FQ_MOD {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}
// r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
#define COZK_FR_MOD {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}
// R mod p (Montgomery one)
#define COZK_FQ_ONE {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}
// R^2 mod p
#define COZK_FQ_R2 {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u}
#define COZK_FQ_N0 0xe4866389u  // -p^-1 mod 2^32
#define COZK_FR_N0 0xefffffffu  // -r^-1 mod 2^32

#if defined(__CUDA_ARCH__)
#include "field_ptx.inc"
#if defined(COZK_FIELD_CALLS)
// The three multiplication forms as real functions (register ABI: arguments and result travel in registers, no stack
// frame).  A translation unit defines COZK_FIELD_CALLS before its first include when its kernels are bound by depth on a
// few warps and should stay small (depth_kernels.cu); everywhere else the operations expand inline.
namespace cozk_calls {
struct f8 {
    uint32_t v[8];
};
static __device__ __noinline__ f8 mul(f8 a, f8 b) {
    f8 r;
    fq_mul_ptx(r.v, a.v, b.v);
    return r;
}
static __device__ __noinline__ f8 sqr(f8 a) {
    f8 r;
    fq_sqr_ptx(r.v, a.v);
    return r;
}
static __device__ __noinline__ f8 mul2(f8 a, f8 b, f8 c, f8 d) {
    f8 r;
    fq_mul2_ptx(r.v, a.v, b.v, c.v, d.v);
    return r;
}
}  // namespace cozk_calls
#endif
#endif

// ------------------------------------------------------------------ portable (host) bodies
namespace host {
inline bool geq(const uint32_t* a, const uint32_t* m) {
    for (int i = 7; i >= 0; --i) {
        if (a[i] > m[i]) return true;
        if (a[i] < m[i]) return false;
    }
    return true;
}
inline void sub_raw(uint32_t* a, const uint32_t* m) {
    uint64_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a[i] - m[i] - br;
        a[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
}
#if defined(__SIZEOF_INT128__)
// 4 x 64-bit limbs with unsigned __int128: the host runs the serial tail of an MSM (Horner over 254 bit positions,
// one inversion) an order of magnitude faster than a single GPU thread, so this body is tuned for latency.
inline void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* mod, uint32_t n0) {
    typedef unsigned __int128 u128;
    const uint64_t n064 = (n0 == 0xe4866389u) ? 0x87d20782e4866389ULL : 0xc2e1f593efffffffULL;
    uint64_t A[4], B[4], M[4];
    for (int i = 0; i < 4; ++i) {
        A[i] = a[2 * i] | ((uint64_t)a[2 * i + 1] << 32);
        B[i] = b[2 * i] | ((uint64_t)b[2 * i + 1] << 32);
        M[i] = mod[2 * i] | ((uint64_t)mod[2 * i + 1] << 32);
    }
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    for (int i = 0; i < 4; ++i) {
        u128 c;
        c = (u128)A[0] * B[i] + t0; t0 = (uint64_t)c; c >>= 64;
        c += (u128)A[1] * B[i] + t1; t1 = (uint64_t)c; c >>= 64;
        c += (u128)A[2] * B[i] + t2; t2 = (uint64_t)c; c >>= 64;
        c += (u128)A[3] * B[i] + t3; t3 = (uint64_t)c; c >>= 64;
        c += t4; t4 = (uint64_t)c;
        uint64_t t5 = (uint64_t)(c >> 64);
        uint64_t m = t0 * n064;
        c = (u128)m * M[0] + t0; c >>= 64;
        c += (u128)m * M[1] + t1; t0 = (uint64_t)c; c >>= 64;
        c += (u128)m * M[2] + t2; t1 = (uint64_t)c; c >>= 64;
        c += (u128)m * M[3] + t3; t2 = (uint64_t)c; c >>= 64;
        c += t4; t3 = (uint64_t)c; t4 = t5 + (uint64_t)(c >> 64);
    }
    uint32_t t[8] = {(uint32_t)t0, (uint32_t)(t0 >> 32), (uint32_t)t1, (uint32_t)(t1 >> 32),
                     (uint32_t)t2, (uint32_t)(t2 >> 32), (uint32_t)t3, (uint32_t)(t3 >> 32)};
    if (t4 || geq(t, mod)) sub_raw(t, mod);
    for (int i = 0; i < 8; ++i) r[i] = t[i];
}
#else
inline void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* mod, uint32_t n0) {
    uint32_t t[10] = {0};
    for (int i = 0; i < 8; ++i) {
        uint64_t c = 0;
        for (int j = 0; j < 8; ++j) {
            c += (uint64_t)a[j] * b[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * n0;
        c = (uint64_t)m * mod[0] + t[0];
        c >>= 32;
        for (int j = 1; j < 8; ++j) {
            c += (uint64_t)m * mod[j] + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    if (t[8] || geq(t, mod)) sub_raw(t, mod);
    for (int i = 0; i < 8; ++i) r[i] = t[i];
}
#endif
inline void add_mod(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* mod) {
    uint32_t t[8];
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a[i] + b[i];
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    if (geq(t, mod)) sub_raw(t, mod);
    for (int i = 0; i < 8; ++i) r[i] = t[i];
}
inline void sub_mod(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* mod) {
    uint32_t t[8];
    uint64_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        t[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    if (br) {
        uint64_t c = 0;
        for (int i = 0; i < 8; ++i) {
            c += (uint64_t)t[i] + mod[i];
            t[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    for (int i = 0; i < 8; ++i) r[i] = t[i];
}
}  // namespace host

// ------------------------------------------------------------------ Fq operations (fully reduced in, fully reduced out)
COZK_HD fq fq_mul(const fq& a, const fq& b) {
    fq r;
#if defined(__CUDA_ARCH__) && defined(COZK_FIELD_CALLS)
    cozk_calls::f8 x, y;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x.v[i] = a.v[i];
        y.v[i] = b.v[i];
    }
    const cozk_calls::f8 z = cozk_calls::mul(x, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = z.v[i];
#elif defined(__CUDA_ARCH__)
    fq_mul_ptx(r.v, a.v, b.v);
#else
    const uint32_t mod[8] = COZK_FQ_MOD;
    host::mont_mul(r.v, a.v, b.v, mod, COZK_FQ_N0);
#endif
    return r;
}
COZK_HD fq fq_sqr(const fq& a) {
    fq r;
#if defined(__CUDA_ARCH__) && defined(COZK_FIELD_CALLS)
    cozk_calls::f8 x;
#pragma unroll
    for (int i = 0; i < 8; ++i) x.v[i] = a.v[i];
    const cozk_calls::f8 z = cozk_calls::sqr(x);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = z.v[i];
#elif defined(__CUDA_ARCH__)
    fq_sqr_ptx(r.v, a.v);
#else
    const uint32_t mod[8] = COZK_FQ_MOD;
    host::mont_mul(r.v, a.v, a.v, mod, COZK_FQ_N0);
#endif
    return r;
}
COZK_HD fq fq_add(const fq& a, const fq& b);
// a*b + c*d with ONE Montgomery reduction (200 multiply-adds instead of 272); same fully reduced value as
// fq_add(fq_mul(a, b), fq_mul(c, d)), which is what the host build computes
COZK_HD fq fq_mul2(const fq& a, const fq& b, const fq& c, const fq& d) {
#if defined(__CUDA_ARCH__) && defined(COZK_FIELD_CALLS)
    cozk_calls::f8 x, y, z, w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x.v[i] = a.v[i];
        y.v[i] = b.v[i];
        z.v[i] = c.v[i];
        w.v[i] = d.v[i];
    }
    const cozk_calls::f8 o = cozk_calls::mul2(x, y, z, w);
    fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = o.v[i];
    return r;
#elif defined(__CUDA_ARCH__)
    fq r;
    fq_mul2_ptx(r.v, a.v, b.v, c.v, d.v);
    return r;
#else
    return fq_add(fq_mul(a, b), fq_mul(c, d));
#endif
}
COZK_HD fq fq_add(const fq& a, const fq& b) {
    fq r;
#if defined(__CUDA_ARCH__)
    fq_add_ptx(r.v, a.v, b.v);
#else
    const uint32_t mod[8] = COZK_FQ_MOD;
    host::add_mod(r.v, a.v, b.v, mod);
#endif
    return r;
}
COZK_HD fq fq_sub(const fq& a, const fq& b) {
    fq r;
#if defined(__CUDA_ARCH__)
    fq_sub_ptx(r.v, a.v, b.v);
#else
    const uint32_t mod[8] = COZK_FQ_MOD;
    host::sub_mod(r.v, a.v, b.v, mod);
#endif
    return r;
}
COZK_HD fq fq_zero() {
    fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0;
    return r;
}
COZK_HD fq fq_one() {
    const uint32_t one[8] = COZK_FQ_ONE;
    fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = one[i];
    return r;
}
COZK_HD bool fq_is_zero(const fq& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.v[i];
    return o == 0;
}
COZK_HD bool fq_eq(const fq& a, const fq& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
COZK_HD fq fq_neg(const fq& a) { return fq_sub(fq_zero(), a); }
COZK_HD fq fq_dbl(const fq& a) { return fq_add(a, a); }
// y := cond ? -y : y   (y = 0 stays 0)
COZK_HD fq fq_cneg(const fq& a, bool cond) {
    fq n = fq_neg(a), r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = cond ? n.v[i] : a.v[i];
    return r;
}
// a^(p-2): Fermat inversion, square-and-multiply over the fixed exponent; 0 -> 0.  Used once per MSM result.
COZK_HD fq fq_inv(const fq& a) {
    const uint32_t e[8] = {0xd87cfd45u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    fq acc = fq_one();
    for (int i = 253; i >= 0; --i) {
        acc = fq_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fq_mul(acc, a);
    }
    return acc;
}

// ------------------------------------------------------------------ Fr: only what scalar decoding needs
COZK_HD fr fr_mul(const fr& a, const fr& b) {
    fr r;
#if defined(__CUDA_ARCH__)
    fr_mul_ptx(r.v, a.v, b.v);
#else
    const uint32_t mod[8] = COZK_FR_MOD;
    host::mont_mul(r.v, a.v, b.v, mod, COZK_FR_N0);
#endif
    return r;
}
COZK_HD fr fr_add(const fr& a, const fr& b) {
    fr r;
#if defined(__CUDA_ARCH__)
    fr_add_ptx(r.v, a.v, b.v);
#else
    const uint32_t mod[8] = COZK_FR_MOD;
    host::add_mod(r.v, a.v, b.v, mod);
#endif
    return r;
}
COZK_HD fr fr_sub(const fr& a, const fr& b) {
    fr r;
#if defined(__CUDA_ARCH__)
    fr_sub_ptx(r.v, a.v, b.v);
#else
    const uint32_t mod[8] = COZK_FR_MOD;
    host::sub_mod(r.v, a.v, b.v, mod);
#endif
    return r;
}
// canonical integer of a Montgomery-form scalar: s * R^-1 mod r
COZK_HD fr fr_from_mont(const fr& a) {
    fr one = fq_zero();
    one.v[0] = 1;
    return fr_mul(a, one);
}
// integer in [0, 2^256) -> [0, r): at most five subtractions (2^256 / r < 6)
COZK_HD fr fr_reduce_canon(const fr& a) {
    const uint32_t mod[8] = COZK_FR_MOD;
    fr t = a;
    for (int k = 0; k < 6; ++k) {
        uint32_t s[8];
        uint32_t br = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint64_t d = (uint64_t)t.v[i] - mod[i] - br;
            s[i] = (uint32_t)d;
            br = (uint32_t)(d >> 32) & 1u;
        }
        if (br) break;
#pragma unroll
        for (int i = 0; i < 8; ++i) t.v[i] = s[i];
    }
    return t;
}

}  // namespace cozk
