/* oracle/bn254.c - see oracle/bn254.h.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no reference vectors exist).
 *
 * Independent of the CUDA engine by construction: 4 x 64-bit limbs with unsigned __int128 (the limb layout
 * of ark-ff's Fp256<MontBackend<_, 4>>), Jacobian coordinates (ark-ec short_weierstrass::Projective), and
 * the window/digit/bucket schedule of ark-ec 0.5 scalar_mul::variable_base (restated from its published
 * algorithm; the crate source is not available offline).
 */
#define _GNU_SOURCE
#include "bn254.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct { u64 l[4]; } fe;
typedef struct { u64 mod[4]; u64 n0; fe r1; fe r2; } fparams; /* r1 = R mod m, r2 = R^2 mod m, n0 = -m^-1 mod 2^64 */

static const fparams FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}}};
static const fparams FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}}};

/* ------------------------------------------------------------------------------------------------ field */

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int geq_mod(const u64* a, const u64* m) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > m[i]) return 1;
        if (a[i] < m[i]) return 0;
    }
    return 1;
}
static inline void sub_mod_raw(u64* a, const u64* m) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - m[i] - (u64)br;
        a[i] = (u64)d;
        br = (d >> 64) & 1;
    }
}
#define FE_INLINE static inline __attribute__((always_inline))
/* t (4 limbs, < 2*mod) -> t mod m, branch-free (random operands make a compare-and-branch mispredict half the time) */
FE_INLINE void reduce_once(fe* o, u64 t0, u64 t1, u64 t2, u64 t3, const fparams* F) {
    u128 d;
    u64 s0, s1, s2, s3, br;
    d = (u128)t0 - F->mod[0]; s0 = (u64)d; br = (u64)(d >> 64) & 1;
    d = (u128)t1 - F->mod[1] - br; s1 = (u64)d; br = (u64)(d >> 64) & 1;
    d = (u128)t2 - F->mod[2] - br; s2 = (u64)d; br = (u64)(d >> 64) & 1;
    d = (u128)t3 - F->mod[3] - br; s3 = (u64)d; br = (u64)(d >> 64) & 1;
    u64 keep = (u64)0 - br; /* all ones if t < mod */
    o->l[0] = (t0 & keep) | (s0 & ~keep);
    o->l[1] = (t1 & keep) | (s1 & ~keep);
    o->l[2] = (t2 & keep) | (s2 & ~keep);
    o->l[3] = (t3 & keep) | (s3 & ~keep);
}
FE_INLINE void fe_add(fe* o, const fe* a, const fe* b, const fparams* F) {
    u128 c;
    u64 t0, t1, t2, t3;
    c = (u128)a->l[0] + b->l[0]; t0 = (u64)c; c >>= 64;
    c += (u128)a->l[1] + b->l[1]; t1 = (u64)c; c >>= 64;
    c += (u128)a->l[2] + b->l[2]; t2 = (u64)c; c >>= 64;
    c += (u128)a->l[3] + b->l[3]; t3 = (u64)c; /* both moduli are 254-bit: no carry out of 256 bits */
    reduce_once(o, t0, t1, t2, t3, F);
}
FE_INLINE void fe_sub(fe* o, const fe* a, const fe* b, const fparams* F) {
    u128 d;
    u64 t0, t1, t2, t3, br;
    d = (u128)a->l[0] - b->l[0]; t0 = (u64)d; br = (u64)(d >> 64) & 1;
    d = (u128)a->l[1] - b->l[1] - br; t1 = (u64)d; br = (u64)(d >> 64) & 1;
    d = (u128)a->l[2] - b->l[2] - br; t2 = (u64)d; br = (u64)(d >> 64) & 1;
    d = (u128)a->l[3] - b->l[3] - br; t3 = (u64)d; br = (u64)(d >> 64) & 1;
    u64 mask = (u64)0 - br;
    u128 c;
    c = (u128)t0 + (F->mod[0] & mask); o->l[0] = (u64)c; c >>= 64;
    c += (u128)t1 + (F->mod[1] & mask); o->l[1] = (u64)c; c >>= 64;
    c += (u128)t2 + (F->mod[2] & mask); o->l[2] = (u64)c; c >>= 64;
    c += (u128)t3 + (F->mod[3] & mask); o->l[3] = (u64)c;
}
FE_INLINE void fe_neg(fe* o, const fe* a, const fparams* F) {
    fe z = {{0, 0, 0, 0}};
    fe_sub(o, &z, a, F);
}
FE_INLINE void fe_dbl(fe* o, const fe* a, const fparams* F) { fe_add(o, a, a, F); }

/* CIOS Montgomery product, 4 limbs; the two spare top bits of both moduli keep t[4] a single limb.
 * Two bodies: mulx/adc intrinsics (what a tuned CPU library such as ark-ff's asm backend compiles to; ~75
 * cycles here) and a portable unsigned __int128 one (~96 cycles).  Same results, tested against pyref. */
#if defined(__BMI2__) && defined(__x86_64__)
#include <immintrin.h>
FE_INLINE void fe_mul(fe* o, const fe* a, const fe* b, const fparams* F) {
    typedef unsigned long long ull;
    ull a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    ull m0 = F->mod[0], m1 = F->mod[1], m2 = F->mod[2], m3 = F->mod[3];
    ull t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    for (int i = 0; i < 4; ++i) {
        ull bi = b->l[i], l0, l1, l2, l3, h0, h1, h2, h3;
        unsigned char c;
        l0 = _mulx_u64(a0, bi, &h0); l1 = _mulx_u64(a1, bi, &h1); l2 = _mulx_u64(a2, bi, &h2); l3 = _mulx_u64(a3, bi, &h3);
        c = _addcarry_u64(0, l1, h0, &l1); c = _addcarry_u64(c, l2, h1, &l2); c = _addcarry_u64(c, l3, h2, &l3); _addcarry_u64(c, h3, 0, &h3);
        c = _addcarry_u64(0, t0, l0, &t0); c = _addcarry_u64(c, t1, l1, &t1); c = _addcarry_u64(c, t2, l2, &t2); c = _addcarry_u64(c, t3, l3, &t3);
        _addcarry_u64(c, t4, h3, &t4);
        ull m = t0 * F->n0, dead;
        l0 = _mulx_u64(m, m0, &h0); l1 = _mulx_u64(m, m1, &h1); l2 = _mulx_u64(m, m2, &h2); l3 = _mulx_u64(m, m3, &h3);
        c = _addcarry_u64(0, l1, h0, &l1); c = _addcarry_u64(c, l2, h1, &l2); c = _addcarry_u64(c, l3, h2, &l3); _addcarry_u64(c, h3, 0, &h3);
        c = _addcarry_u64(0, t0, l0, &dead); c = _addcarry_u64(c, t1, l1, &t0); c = _addcarry_u64(c, t2, l2, &t1); c = _addcarry_u64(c, t3, l3, &t2);
        _addcarry_u64(c, t4, h3, &t3);
        t4 = 0; /* a*b + m*mod < 2^256 * 2*mod with 254-bit moduli: the top limb never overflows */
    }
    reduce_once(o, t0, t1, t2, t3, F);
}
#else
FE_INLINE void fe_mul(fe* o, const fe* a, const fe* b, const fparams* F) {
    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    for (int i = 0; i < 4; ++i) {
        u64 bi = b->l[i];
        u128 c;
        c = (u128)a->l[0] * bi + t0; t0 = (u64)c; c >>= 64;
        c += (u128)a->l[1] * bi + t1; t1 = (u64)c; c >>= 64;
        c += (u128)a->l[2] * bi + t2; t2 = (u64)c; c >>= 64;
        c += (u128)a->l[3] * bi + t3; t3 = (u64)c; c >>= 64;
        c += t4; t4 = (u64)c; u64 t5 = (u64)(c >> 64);
        u64 m = t0 * F->n0;
        c = (u128)m * F->mod[0] + t0; c >>= 64;
        c += (u128)m * F->mod[1] + t1; t0 = (u64)c; c >>= 64;
        c += (u128)m * F->mod[2] + t2; t1 = (u64)c; c >>= 64;
        c += (u128)m * F->mod[3] + t3; t2 = (u64)c; c >>= 64;
        c += t4; t3 = (u64)c; t4 = t5 + (u64)(c >> 64);
    }
    fe t = {{t0, t1, t2, t3}};
    if (t4 || geq_mod(t.l, F->mod)) sub_mod_raw(t.l, F->mod);
    *o = t;
}
#endif
FE_INLINE void fe_sqr(fe* o, const fe* a, const fparams* F) { fe_mul(o, a, a, F); }
static void fe_to_mont(fe* o, const fe* a, const fparams* F) { fe_mul(o, a, &F->r2, F); }
static void fe_from_mont(fe* o, const fe* a, const fparams* F) {
    fe one = {{1, 0, 0, 0}};
    fe_mul(o, a, &one, F);
}
/* a^e, e given as 4 limbs; a in Montgomery form */
static void fe_pow(fe* o, const fe* a, const u64* e, const fparams* F) {
    fe acc = F->r1;
    int started = 0;
    for (int i = 255; i >= 0; --i) {
        if (started) fe_sqr(&acc, &acc, F);
        if ((e[i >> 6] >> (i & 63)) & 1) {
            if (started) fe_mul(&acc, &acc, a, F); else { acc = *a; started = 1; }
        }
    }
    *o = acc;
}
static void fe_inv(fe* o, const fe* a, const fparams* F) {
    if (fe_is_zero(a)) { *o = *a; return; }
    u64 e[4];
    memcpy(e, F->mod, 32);
    e[0] -= 2; /* both moduli end in ...47 / ...01: no borrow */
    fe_pow(o, a, e, F);
}

void orc_field_op(int which, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    const fparams* F = which ? &FR : &FQ;
    for (size_t i = 0; i < n; ++i) {
        fe x, y, z;
        memcpy(&x, a + 4 * i, 32);
        if (b) memcpy(&y, b + 4 * i, 32); else memset(&y, 0, 32);
        switch (op) {
            case 0: fe_add(&z, &x, &y, F); break;
            case 1: fe_sub(&z, &x, &y, F); break;
            case 2: fe_mul(&z, &x, &y, F); break;
            case 3: fe_sqr(&z, &x, F); break;
            case 4: fe_neg(&z, &x, F); break;
            case 5: fe_inv(&z, &x, F); break;
            case 6: fe_to_mont(&z, &x, F); break;
            default: fe_from_mont(&z, &x, F); break;
        }
        memcpy(out + 4 * i, &z, 32);
    }
}

/* ------------------------------------------------------------------------------------------------ curve */

typedef struct { fe x, y; } aff;    /* finite affine point, Montgomery coordinates */
typedef struct { fe x, y, z; } jac; /* z == 0 <=> identity */

static __thread u64 cnt_madd, cnt_add, cnt_dbl;

static inline void jac_set_identity(jac* p) { memset(p, 0, sizeof *p); p->x = FQ.r1; p->y = FQ.r1; }
static inline int jac_is_identity(const jac* p) { return fe_is_zero(&p->z); }

/* dbl-2009-l, a = 0 */
static void jac_dbl(jac* o, const jac* p) {
    ++cnt_dbl;
    if (jac_is_identity(p)) { *o = *p; return; }
    const fparams* F = &FQ;
    fe A, B, C, D, E, Fq_, t, X3, Y3, Z3;
    fe_sqr(&A, &p->x, F);
    fe_sqr(&B, &p->y, F);
    fe_sqr(&C, &B, F);
    fe_add(&t, &p->x, &B, F);
    fe_sqr(&t, &t, F);
    fe_sub(&t, &t, &A, F);
    fe_sub(&t, &t, &C, F);
    fe_dbl(&D, &t, F);
    fe_dbl(&E, &A, F);
    fe_add(&E, &E, &A, F);
    fe_sqr(&Fq_, &E, F);
    fe_dbl(&t, &D, F);
    fe_sub(&X3, &Fq_, &t, F);
    fe_sub(&t, &D, &X3, F);
    fe_mul(&Y3, &E, &t, F);
    fe_dbl(&t, &C, F); fe_dbl(&t, &t, F); fe_dbl(&t, &t, F);
    fe_sub(&Y3, &Y3, &t, F);
    fe_mul(&Z3, &p->y, &p->z, F);
    fe_dbl(&Z3, &Z3, F);
    o->x = X3; o->y = Y3; o->z = Z3;
}

/* madd-2007-bl with the exceptional cases arkworks handles (P = Q -> double, P = -Q -> identity) */
static void jac_madd(jac* o, const jac* p, const aff* q) {
    ++cnt_madd;
    const fparams* F = &FQ;
    if (jac_is_identity(p)) { o->x = q->x; o->y = q->y; o->z = F->r1; return; }
    fe Z1Z1, U2, S2, H, HH, I, J, r, V, t, X3, Y3, Z3;
    fe_sqr(&Z1Z1, &p->z, F);
    fe_mul(&U2, &q->x, &Z1Z1, F);
    fe_mul(&S2, &q->y, &p->z, F);
    fe_mul(&S2, &S2, &Z1Z1, F);
    if (fe_eq(&U2, &p->x)) {
        if (fe_eq(&S2, &p->y)) { --cnt_madd; jac_dbl(o, p); return; }
        jac_set_identity(o); o->z = (fe){{0, 0, 0, 0}};
        return;
    }
    fe_sub(&H, &U2, &p->x, F);
    fe_sqr(&HH, &H, F);
    fe_dbl(&I, &HH, F); fe_dbl(&I, &I, F);
    fe_mul(&J, &H, &I, F);
    fe_sub(&r, &S2, &p->y, F); fe_dbl(&r, &r, F);
    fe_mul(&V, &p->x, &I, F);
    fe_sqr(&X3, &r, F);
    fe_sub(&X3, &X3, &J, F);
    fe_dbl(&t, &V, F);
    fe_sub(&X3, &X3, &t, F);
    fe_sub(&t, &V, &X3, F);
    fe_mul(&Y3, &r, &t, F);
    fe_mul(&t, &p->y, &J, F); fe_dbl(&t, &t, F);
    fe_sub(&Y3, &Y3, &t, F);
    fe_add(&Z3, &p->z, &H, F);
    fe_sqr(&Z3, &Z3, F);
    fe_sub(&Z3, &Z3, &Z1Z1, F);
    fe_sub(&Z3, &Z3, &HH, F);
    o->x = X3; o->y = Y3; o->z = Z3;
}

/* add-2007-bl */
static void jac_add(jac* o, const jac* p, const jac* q) {
    ++cnt_add;
    const fparams* F = &FQ;
    if (jac_is_identity(p)) { *o = *q; return; }
    if (jac_is_identity(q)) { *o = *p; return; }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, r, V, t, X3, Y3, Z3;
    fe_sqr(&Z1Z1, &p->z, F);
    fe_sqr(&Z2Z2, &q->z, F);
    fe_mul(&U1, &p->x, &Z2Z2, F);
    fe_mul(&U2, &q->x, &Z1Z1, F);
    fe_mul(&S1, &p->y, &q->z, F); fe_mul(&S1, &S1, &Z2Z2, F);
    fe_mul(&S2, &q->y, &p->z, F); fe_mul(&S2, &S2, &Z1Z1, F);
    if (fe_eq(&U1, &U2)) {
        if (fe_eq(&S1, &S2)) { --cnt_add; jac_dbl(o, p); return; }
        jac_set_identity(o); o->z = (fe){{0, 0, 0, 0}};
        return;
    }
    fe_sub(&H, &U2, &U1, F);
    fe_dbl(&I, &H, F); fe_sqr(&I, &I, F);
    fe_mul(&J, &H, &I, F);
    fe_sub(&r, &S2, &S1, F); fe_dbl(&r, &r, F);
    fe_mul(&V, &U1, &I, F);
    fe_sqr(&X3, &r, F);
    fe_sub(&X3, &X3, &J, F);
    fe_dbl(&t, &V, F);
    fe_sub(&X3, &X3, &t, F);
    fe_sub(&t, &V, &X3, F);
    fe_mul(&Y3, &r, &t, F);
    fe_mul(&t, &S1, &J, F); fe_dbl(&t, &t, F);
    fe_sub(&Y3, &Y3, &t, F);
    fe_add(&Z3, &p->z, &q->z, F);
    fe_sqr(&Z3, &Z3, F);
    fe_sub(&Z3, &Z3, &Z1Z1, F);
    fe_sub(&Z3, &Z3, &Z2Z2, F);
    fe_mul(&Z3, &Z3, &H, F);
    o->x = X3; o->y = Y3; o->z = Z3;
}

static inline void aff_neg(aff* o, const aff* p) { o->x = p->x; fe_neg(&o->y, &p->y, &FQ); }

/* 72-byte wire point <-> internal */
static int load72(const uint8_t* in, aff* a) { /* returns 1 if identity */
    memcpy(&a->x, in, 32);
    memcpy(&a->y, in + 32, 32);
    return in[64] != 0;
}
static void store72_jac(const jac* p, uint8_t* out) {
    memset(out, 0, 72);
    if (jac_is_identity(p)) { out[64] = 1; return; }
    const fparams* F = &FQ;
    fe zi, zi2, zi3, x, y;
    fe_inv(&zi, &p->z, F);
    fe_sqr(&zi2, &zi, F);
    fe_mul(&zi3, &zi2, &zi, F);
    fe_mul(&x, &p->x, &zi2, F);
    fe_mul(&y, &p->y, &zi3, F);
    memcpy(out, &x, 32);
    memcpy(out + 32, &y, 32);
}
static void jac_from72(const uint8_t* in, jac* p) {
    aff a;
    if (load72(in, &a)) { jac_set_identity(p); p->z = (fe){{0, 0, 0, 0}}; return; }
    p->x = a.x; p->y = a.y; p->z = FQ.r1;
}

void orc_g1_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        jac pa, pb, r;
        jac_from72(a + 72 * i, &pa);
        if (op == 0) {
            jac_from72(b + 72 * i, &pb);
            jac_add(&r, &pa, &pb);
        } else if (op == 1) {
            jac_dbl(&r, &pa);
        } else {
            r = pa;
            fe_neg(&r.y, &pa.y, &FQ);
        }
        store72_jac(&r, out + 72 * i);
    }
}

int orc_g1_is_valid(const uint8_t* pt) {
    aff a;
    if (load72(pt, &a)) return 1;
    if (geq_mod(a.x.l, FQ.mod) || geq_mod(a.y.l, FQ.mod)) return 0;
    fe y2, x3, three, t;
    fe_sqr(&y2, &a.y, &FQ);
    fe_sqr(&x3, &a.x, &FQ);
    fe_mul(&x3, &x3, &a.x, &FQ);
    fe_add(&t, &FQ.r1, &FQ.r1, &FQ);
    fe_add(&three, &t, &FQ.r1, &FQ);
    fe_add(&x3, &x3, &three, &FQ);
    return fe_eq(&y2, &x3);
}

/* canonical 4-limb integer of scalar i */
static void load_scalar(const uint8_t* scalars, size_t stride, int form, size_t i, u64* out) {
    fe s;
    memcpy(&s, scalars + i * stride, 32);
    if (form == ORC_MONT) fe_from_mont(&s, &s, &FR);
    memcpy(out, &s, 32);
}

static void jac_mul_bits(jac* o, const aff* base, const u64* k) {
    jac acc;
    jac_set_identity(&acc); acc.z = (fe){{0, 0, 0, 0}};
    for (int i = 255; i >= 0; --i) {
        jac_dbl(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) jac_madd(&acc, &acc, base);
    }
    *o = acc;
}

void orc_g1_mul(const uint8_t* pt72, const uint64_t* scalar, int form, uint8_t* out72) {
    aff a;
    jac r;
    if (load72(pt72, &a)) { memset(out72, 0, 72); out72[64] = 1; return; }
    u64 k[4];
    load_scalar((const uint8_t*)scalar, 32, form, 0, k);
    jac_mul_bits(&r, &a, k);
    store72_jac(&r, out72);
}

/* ------------------------------------------------------------------------------------------------ generators */

static inline u64 mix64(u64 z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline u64 gen_limb(u64 seed, u64 i, u64 j) { return mix64(mix64(seed) + 4 * i + j); }
static void raw254(u64 seed, u64 i, const u64* mod, u64* out) {
    for (int j = 0; j < 4; ++j) out[j] = gen_limb(seed, i, (u64)j);
    out[3] &= 0x3FFFFFFFFFFFFFFFULL;
    if (geq_mod(out, mod)) sub_mod_raw(out, mod);
}

static void gen_one_base(u64 seed, u64 i, uint8_t* out64) {
    const fparams* F = &FQ;
    static const u64 EXP[4] = {0x4f082305b61f3f52ULL, 0x65e05aa45a1c72a3ULL, 0x6e14116da0605617ULL, 0x0c19139cb84c680aULL}; /* (p+1)/4 */
    u64 raw[4];
    raw254(seed, i, F->mod, raw);
    fe x, three, t, rhs, y, y2;
    memcpy(&x, raw, 32);
    fe_to_mont(&x, &x, F);
    fe_add(&t, &F->r1, &F->r1, F);
    fe_add(&three, &t, &F->r1, F);
    for (;;) {
        fe_sqr(&rhs, &x, F);
        fe_mul(&rhs, &rhs, &x, F);
        fe_add(&rhs, &rhs, &three, F);
        fe_pow(&y, &rhs, EXP, F);
        fe_sqr(&y2, &y, F);
        if (fe_eq(&y2, &rhs)) break;
        fe_add(&x, &x, &F->r1, F);
    }
    fe yc;
    fe_from_mont(&yc, &y, F);
    if (yc.l[0] & 1) fe_neg(&y, &y, F);
    memcpy(out64, &x, 32);
    memcpy(out64 + 32, &y, 32);
}

typedef struct { u64 seed; size_t start, n; uint8_t* out; int tid, nthreads; } gen_job;
static void* gen_worker(void* arg) {
    gen_job* j = (gen_job*)arg;
    for (size_t i = (size_t)j->tid; i < j->n; i += (size_t)j->nthreads) gen_one_base(j->seed, j->start + i, j->out + 64 * i);
    return NULL;
}
void orc_gen_bases(uint64_t seed, size_t start, size_t n, uint8_t* out64, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    gen_job jobs[256];
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (gen_job){seed, start, n, out64, t, threads};
        pthread_create(&th[t], NULL, gen_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}

void orc_gen_scalars(int dist, uint64_t seed, size_t start, size_t n, size_t total_n, int form, uint8_t* out, size_t stride) {
    const fparams* F = &FR;
    u64 c0[4], c1[4];
    raw254(seed, 0, F->mod, c0);
    raw254(seed, 1, F->mod, c1);
    for (size_t k = 0; k < n; ++k) {
        size_t i = start + k;
        fe s = {{0, 0, 0, 0}};
        switch (dist) {
            case ORC_DIST_UNIFORM: raw254(seed, i, F->mod, s.l); break;
            case ORC_DIST_CONST: memcpy(s.l, c0, 32); break;
            case ORC_DIST_WMINUS: {
                fe w = {{gen_limb(seed, i + 2, 0) & 0xFFFFFFFFULL, 0, 0, 0}}, a, b;
                memcpy(a.l, c0, 32);
                memcpy(b.l, c1, 32);
                fe_sub(&s, &w, &a, F);
                fe_sub(&s, &s, &b, F);
                break;
            }
            case ORC_DIST_DUP: raw254(seed, i >> 1, F->mod, s.l); break;
            case ORC_DIST_SMALL16: s.l[0] = gen_limb(seed, i, 0) & 0xFFFF; break;
            default: if (i < (total_n + 1) / 2) raw254(seed, i, F->mod, s.l); break;
        }
        if (form == ORC_MONT) fe_to_mont(&s, &s, F);
        memcpy(out + k * stride, &s, 32);
    }
}

/* ------------------------------------------------------------------------------------------------ MSM */

static void load_base(const uint8_t* bases64, size_t i, aff* a) {
    memcpy(&a->x, bases64 + 64 * i, 32);
    memcpy(&a->y, bases64 + 64 * i + 32, 32);
}

void orc_msm_naive(const uint8_t* bases64, const uint8_t* scalars, size_t stride, int form, size_t n, uint8_t* out72) {
    jac acc;
    jac_set_identity(&acc); acc.z = (fe){{0, 0, 0, 0}};
    for (size_t i = 0; i < n; ++i) {
        aff b;
        u64 k[4];
        jac t;
        load_base(bases64, i, &b);
        load_scalar(scalars, stride, form, i, k);
        jac_mul_bits(&t, &b, k);
        jac_add(&acc, &acc, &t);
    }
    store72_jac(&acc, out72);
}

/* ark_std::log2 = ceil(log2(x)) for x > 1, 0 for x <= 1 */
static unsigned ceil_log2(size_t x) {
    unsigned l = 0;
    while (((size_t)1 << l) < x) ++l;
    return l;
}
int orc_msm_window(size_t n) { return n < 32 ? 3 : (int)(ceil_log2(n) * 69 / 100) + 2; }

#define SCALAR_BITS 254

/* ark-ec make_digits: signed base-2^w digits, each in [-2^(w-1), 2^(w-1)); the last digit absorbs the final carry */
static void make_digits(const u64* a, unsigned w, int32_t* digits, unsigned count) {
    u64 radix = (u64)1 << w, mask = radix - 1, carry = 0;
    for (unsigned i = 0; i < count; ++i) {
        unsigned off = i * w, idx = off / 64, bit = off % 64;
        u64 buf;
        if (bit < 64 - w || idx == 3) buf = a[idx] >> bit;
        else buf = (a[idx] >> bit) | (a[idx + 1] << (64 - bit));
        u64 coef = carry + (buf & mask);
        carry = (coef + radix / 2) >> w;
        digits[i] = (int32_t)((int64_t)coef - (int64_t)(carry << w));
    }
    digits[count - 1] += (int32_t)(carry << w);
}

/* digits of the last window are not re-centred: they reach 2^(bits left in the scalar) (top bits all set, plus carry) */
static size_t last_window_buckets(unsigned c, unsigned count) {
    unsigned bits_last = SCALAR_BITS - (count - 1) * c;
    if (bits_last > c) bits_last = c;
    return ((size_t)1 << bits_last) + 1;
}

typedef struct {
    const uint8_t* bases;
    const int32_t* digits; /* n x count */
    size_t n;
    unsigned c, count, nchunks;
    jac* results; /* count x nchunks */
    volatile long next;
    u64 madd, add, dbl;
    pthread_mutex_t mu;
} msm_job;

static void window_chunk_sum(const msm_job* J, unsigned win, size_t lo, size_t hi, jac* out) {
    size_t nb = (size_t)1 << (J->c - 1);
    /* the last window holds an un-recentred digit (see make_digits): the scalar's remaining top bits plus the carry */
    if (win == J->count - 1) nb = last_window_buckets(J->c, J->count);
    jac* buckets = (jac*)calloc(nb, sizeof(jac));
    for (size_t i = lo; i < hi; ++i) {
        int32_t d = J->digits[i * J->count + win];
        if (d == 0) continue;
        aff b;
        load_base(J->bases, i, &b);
        if (d > 0) {
            jac_madd(&buckets[d - 1], &buckets[d - 1], &b);
        } else {
            aff nbp;
            aff_neg(&nbp, &b);
            jac_madd(&buckets[-d - 1], &buckets[-d - 1], &nbp);
        }
    }
    jac run, res;
    memset(&run, 0, sizeof run);
    memset(&res, 0, sizeof res);
    for (size_t k = nb; k-- > 0;) {
        jac_add(&run, &run, &buckets[k]);
        jac_add(&res, &res, &run);
    }
    free(buckets);
    *out = res;
}

typedef struct {
    const uint8_t* scalars;
    size_t stride;
    int form;
    size_t lo, hi;
    unsigned c, count;
    int32_t* digits;
} digit_job;
static void* digit_worker(void* arg) {
    digit_job* d = (digit_job*)arg;
    for (size_t i = d->lo; i < d->hi; ++i) {
        u64 k[4];
        load_scalar(d->scalars, d->stride, d->form, i, k);
        make_digits(k, d->c, d->digits + i * d->count, d->count);
    }
    return NULL;
}

static void* msm_worker(void* arg) {
    msm_job* J = (msm_job*)arg;
    cnt_madd = cnt_add = cnt_dbl = 0;
    long total = (long)J->count * (long)J->nchunks;
    for (;;) {
        long t = __sync_fetch_and_add(&J->next, 1);
        if (t >= total) break;
        unsigned win = (unsigned)(t / J->nchunks), ch = (unsigned)(t % J->nchunks);
        size_t per = (J->n + J->nchunks - 1) / J->nchunks;
        size_t lo = per * ch, hi = lo + per > J->n ? J->n : lo + per;
        if (lo > hi) lo = hi;
        window_chunk_sum(J, win, lo, hi, &J->results[(size_t)win * J->nchunks + ch]);
    }
    pthread_mutex_lock(&J->mu);
    J->madd += cnt_madd; J->add += cnt_add; J->dbl += cnt_dbl;
    pthread_mutex_unlock(&J->mu);
    return NULL;
}

static __thread u64 last_madd, last_add, last_dbl;
void orc_msm_last_counts(uint64_t* madd, uint64_t* add, uint64_t* dbl) { *madd = last_madd; *add = last_add; *dbl = last_dbl; }

void orc_msm(const uint8_t* bases64, const uint8_t* scalars, size_t stride, int form, size_t n, int threads, uint8_t* out72) {
    if (n == 0) { memset(out72, 0, 72); out72[64] = 1; return; }
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    unsigned c = (unsigned)orc_msm_window(n);
    unsigned count = (SCALAR_BITS + c - 1) / c;
    int32_t* digits = (int32_t*)malloc(n * count * sizeof(int32_t));
    /* digit recoding of all scalars, spread over the threads like everything else */
    {
        digit_job dj[512];
        pthread_t th[512];
        int nt = threads;
        if ((size_t)nt > n) nt = (int)n;
        for (int t = 0; t < nt; ++t) {
            dj[t] = (digit_job){scalars, stride, form, n * (size_t)t / (size_t)nt, n * (size_t)(t + 1) / (size_t)nt, c, count, digits};
            if (nt > 1) pthread_create(&th[t], NULL, digit_worker, &dj[t]);
        }
        if (nt == 1) digit_worker(&dj[0]);
        else for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    }
    msm_job J;
    memset(&J, 0, sizeof J);
    J.bases = bases64; J.digits = digits; J.n = n; J.c = c; J.count = count;
    /* arkworks spreads one MSM over its windows only (<= 17 threads busy, two rounds on 16 cores); jolt-core's batch_msm
       fills the cores with whole polynomials.  Here: window x point-chunk tasks, the chunk count chosen to minimise
       rounds x (points per chunk + the bucket reduction every chunk repeats). */
    J.nchunks = 1;
    if (threads > 1) {
        double best = 0;
        for (unsigned nc = 1; nc <= 32; ++nc) {
            double tasks = (double)count * nc, rounds = (double)(size_t)((tasks + threads - 1) / threads);
            double cost = rounds * ((double)n / nc + 3.0 * (double)((size_t)1 << (c - 1)));
            if (nc == 1 || cost < best) { best = cost; J.nchunks = nc; }
        }
    }
    if ((size_t)J.nchunks > n) J.nchunks = (unsigned)n;
    J.results = (jac*)calloc((size_t)count * J.nchunks, sizeof(jac));
    pthread_mutex_init(&J.mu, NULL);
    if (threads == 1) {
        msm_worker(&J);
    } else {
        pthread_t th[512];
        for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, msm_worker, &J);
        for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    }
    cnt_madd = cnt_add = cnt_dbl = 0;
    /* per-window sums over chunks, then lowest + fold(high..1: total += w; c doublings) */
    jac* wsum = (jac*)calloc(count, sizeof(jac));
    for (unsigned w = 0; w < count; ++w)
        for (unsigned ch = 0; ch < J.nchunks; ++ch) jac_add(&wsum[w], &wsum[w], &J.results[(size_t)w * J.nchunks + ch]);
    jac total;
    memset(&total, 0, sizeof total);
    for (unsigned w = count - 1; w >= 1; --w) {
        jac_add(&total, &total, &wsum[w]);
        for (unsigned k = 0; k < c; ++k) jac_dbl(&total, &total);
    }
    jac_add(&total, &total, &wsum[0]);
    last_madd = J.madd + cnt_madd; last_add = J.add + cnt_add; last_dbl = J.dbl + cnt_dbl;
    store72_jac(&total, out72);
    free(wsum); free(J.results); free(digits);
    pthread_mutex_destroy(&J.mu);
}
