// Thread bodies for the steps on either side of the MSM: they keep a party's
// share device-resident from the moment it arrives until the opening proof leaves.
//
//   N2  ingest         ark-serialize wire image of Rep3DensePolynomial coefficients (32-byte little-endian CANONICAL
//                      integers, a || b per coefficient) -> the in-memory image (Fr Montgomery) the MSM reads at stride 64.
//                      Reference: `receive_request` -> `deserialize_uncompressed_unchecked`
//                      (mpc-net/src/rep3/quic/worker.rs:206-219), struct layout co-jolt/src/poly/dense_mlpoly.rs:23-32,
//                      share layout mpc-types/src/protocols/rep3/arithmetic/types.rs:22-29.
//   N4  linear combination   joint[i] = sum_j coeff_j * poly_j[i] over shared (both halves) and public polynomials, the
//                      public terms added to share a on party 0 and share b on party 1
//                      (co-jolt/src/poly/multilinear_polynomial.rs:196-296; Rep3DensePolynomial::linear_combination,
//                      co-jolt/src/poly/dense_mlpoly.rs:186-224), and chi dot products
//                      (evaluate_at_chi, dense_mlpoly.rs:160-181: sum_i (a_i + b_i)/2 * chi_i, an additive share).
//   N1  pair sums      S[b] = P[2b] + P[2b+1]: open() feeds every quotient scalar to two adjacent bases
//                      (co-jolt/src/poly/commitment/pst13.rs:459), so its MSM equals a half-size MSM over pair sums.
//
//   eq table           chi[b] = prod_i (bit_i(b) ? t_i : 1 - t_i): the table a multilinear evaluation is a dot product with.
//                      co-spartan evaluates every opened polynomial at the opening point
//                      (`p.evaluate(point)` in distributed_batch_open_poly_worker,
//                      co-noir-spartan/co-spartan/src/worker.rs:761-764: ark-poly fixes variable i = bit i of the index
//                      with point[i]); co-jolt builds the same table with the point reversed (EqPolynomial::evals of
//                      jolt-core, the `chis` of evaluate_at_chi).  Built on the device, so the chi table is no longer an
//                      H2D copy of 32 B per coefficient.
//
// Like msm_kernels.cuh, every body is a pure function of its thread index and compiles for the host emulation tier.
//
// Lazy reduction: a sum of products is accumulated as a plain 576-bit integer (64 limb products per term instead of
// the 136 of a Montgomery multiplication) and reduced once per output element.
#pragma once
#include "msm_kernels.cuh"

namespace cozk {

// R^2 mod r (to Montgomery form: fr_mul(x, R2) = x * R)
#define COZK_FR_R2 {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}
// (r + 1) / 2, Montgomery form: the TWO_INV of snarks-core/src/field.rs:6 as the kernels need it
#define COZK_FR_TWO_INV_MONT {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x1f37631au}

COZK_HD bool fr_is_canonical(const fr& a) {
    const uint32_t mod[8] = COZK_FR_MOD;
    uint32_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a.v[i] - mod[i] - br;
        br = (uint32_t)(d >> 32) & 1u;
    }
    return br != 0;  // a < r
}
COZK_HD fr fr_const(const uint32_t (&c)[8]) {
    fr r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = c[i];
    return r;
}
COZK_HD fr fr_mont_from_canon(const fr& a) {
    const uint32_t r2[8] = COZK_FR_R2;
    return fr_mul(a, fr_const(r2));
}
COZK_HD fr fr_neg(const fr& a) { return fr_sub(fq_zero(), a); }
// R mod r (Montgomery one)
#define COZK_FR_ONE {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}
COZK_HD fr fr_one() {
    const uint32_t one[8] = COZK_FR_ONE;
    return fr_const(one);
}

// ------------------------------------------------------------------------------------------------ lazy accumulator
struct fr_wide {
    uint32_t v[18];
};
COZK_HD fr_wide fr_wide_zero() {
    fr_wide w;
#pragma unroll
    for (int i = 0; i < 18; ++i) w.v[i] = 0;
    return w;
}
// acc += a * b  (plain integers; a, b < 2^256; the caller keeps the number of terms below 2^64)
COZK_HD void fr_wide_mac(fr_wide& acc, const fr& a, const fr& b) {
    uint32_t t[16];
#if defined(__CUDA_ARCH__)
    mulwide_ptx(t, a.v, b.v);
    // one 18-limb carry chain (IADD3 / IADD3.X); written as PTX so that the carries stay in the flag
    asm("add.cc.u32 %0, %0, %18;\n\t"
        "addc.cc.u32 %1, %1, %19;\n\t"
        "addc.cc.u32 %2, %2, %20;\n\t"
        "addc.cc.u32 %3, %3, %21;\n\t"
        "addc.cc.u32 %4, %4, %22;\n\t"
        "addc.cc.u32 %5, %5, %23;\n\t"
        "addc.cc.u32 %6, %6, %24;\n\t"
        "addc.cc.u32 %7, %7, %25;\n\t"
        "addc.cc.u32 %8, %8, %26;\n\t"
        "addc.cc.u32 %9, %9, %27;\n\t"
        "addc.cc.u32 %10, %10, %28;\n\t"
        "addc.cc.u32 %11, %11, %29;\n\t"
        "addc.cc.u32 %12, %12, %30;\n\t"
        "addc.cc.u32 %13, %13, %31;\n\t"
        "addc.cc.u32 %14, %14, %32;\n\t"
        "addc.cc.u32 %15, %15, %33;\n\t"
        "addc.cc.u32 %16, %16, 0;\n\t"
        "addc.u32 %17, %17, 0;\n\t"
        : "+r"(acc.v[0]), "+r"(acc.v[1]), "+r"(acc.v[2]), "+r"(acc.v[3]), "+r"(acc.v[4]), "+r"(acc.v[5]), "+r"(acc.v[6]),
          "+r"(acc.v[7]), "+r"(acc.v[8]), "+r"(acc.v[9]), "+r"(acc.v[10]), "+r"(acc.v[11]), "+r"(acc.v[12]), "+r"(acc.v[13]),
          "+r"(acc.v[14]), "+r"(acc.v[15]), "+r"(acc.v[16]), "+r"(acc.v[17])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]), "r"(t[9]),
          "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]));
#else
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t c = 0;
        for (int j = 0; j < 8; ++j) {
            c += (uint64_t)a.v[j] * b.v[i] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
    uint64_t c = 0;
    for (int i = 0; i < 16; ++i) {
        c += (uint64_t)acc.v[i] + t[i];
        acc.v[i] = (uint32_t)c;
        c >>= 32;
    }
    c += acc.v[16];
    acc.v[16] = (uint32_t)c;
    acc.v[17] += (uint32_t)(c >> 32);
#endif
}
// acc * R^-1 mod r, fully reduced.  With acc = lo + hi * 2^256 + top * 2^512:
//     acc / R = lo / R + hi + top * 2^256  (mod r)  =  from_mont(lo) + hi + to_mont(top)
COZK_HD fr fr_wide_reduce(const fr_wide& acc) {
    fr lo, hi, top = fq_zero();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        lo.v[i] = acc.v[i];
        hi.v[i] = acc.v[8 + i];
    }
    top.v[0] = acc.v[16];
    top.v[1] = acc.v[17];
    fr r = fr_add(fr_from_mont(fr_reduce_canon(lo)), fr_reduce_canon(hi));
    if (top.v[0] | top.v[1]) r = fr_add(r, fr_mont_from_canon(top));
    return r;
}

// ------------------------------------------------------------------------------------------------ N2 ingest
struct IngestArgs {
    uint8_t* data;   // n_fr field elements, 32 B each, canonical little-endian -> Montgomery, in place
    size_t n_fr;
    uint32_t* bad;   // set to 1 when an element is >= r (ark-serialize rejects it: SerializationError::InvalidData)
};
COZK_HD void ingest_body(size_t t, const IngestArgs& A) {
    if (t >= A.n_fr) return;
    fr x = load_fq(A.data + 32 * t);
    if (!fr_is_canonical(x)) {
        *A.bad = 1;  // every writer stores the same value
        return;
    }
    store_fq(A.data + 32 * t, fr_mont_from_canon(x));
}

// small public scalars (MultilinearPolynomial::{U8,U16,U32,U64,I64}Scalars, co-jolt/src/poly/multilinear_polynomial.rs:226-268)
// -> 32-byte canonical integers; a negative i64 becomes r - |v| (what F::from_i64 gives)
struct WidenArgs {
    const uint8_t* src;
    uint32_t elem_bytes;  // 1, 2, 4, 8
    uint32_t is_signed;   // only with elem_bytes == 8
    size_t n;
    uint8_t* dst;         // n x 32 B
};
COZK_HD void widen_body(size_t t, const WidenArgs& A) {
    if (t >= A.n) return;
    uint64_t v = 0;
    for (uint32_t k = 0; k < A.elem_bytes; ++k) v |= (uint64_t)A.src[t * A.elem_bytes + k] << (8 * k);
    fr x = fq_zero();
    bool neg = A.is_signed && (v >> 63);
    if (neg) v = (uint64_t)0 - v;
    x.v[0] = (uint32_t)v;
    x.v[1] = (uint32_t)(v >> 32);
    if (neg) {
        const uint32_t mod[8] = COZK_FR_MOD;
        uint32_t br = 0;
        fr y;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint64_t d = (uint64_t)mod[i] - x.v[i] - br;
            y.v[i] = (uint32_t)d;
            br = (uint32_t)(d >> 32) & 1u;
        }
        x = y;
    }
    store_fq(A.dst + 32 * t, x);
}

// ------------------------------------------------------------------------------------------------ N4 linear combination
constexpr uint32_t POLY_SHARED = 0;  // Rep3PrimeFieldShare{a, b}: 64 B per coefficient, Fr Montgomery
constexpr uint32_t POLY_MONT = 1;    // public LargeScalars: 32 B per coefficient, Fr Montgomery
constexpr uint32_t POLY_CANON = 2;   // public small scalars widened to 32-byte canonical integers

struct PolyDesc {
    const uint8_t* data;  // coefficient 0 of the polynomial's chunk_range
    uint64_t len;
    uint32_t kind;
    uint32_t pad;
};
struct LincombArgs {
    const PolyDesc* polys;  // k descriptors
    const fr* coeffs;       // k x 2: [2j] = coeff_j (Montgomery), [2j+1] = coeff_j * R = fr_mont_from_canon(coeff_j): times a CANONICAL value it gives the Montgomery product
    uint32_t k;
    uint32_t party;         // PartyID 0, 1, 2
    uint32_t shared_out;    // 1: out is AoS shares (64 B), 0: every input is public, out is dense Montgomery (32 B)
    size_t n;               // max length
    uint8_t* out;
};
// One term of element i: the two field elements it contributes (x = a / the public value, y = b) and where they go.
struct LincombTerm {
    fr x, y;
    uint32_t kind;  // POLY_* or 0xFFFFFFFF when polynomial j does not reach index i
};
COZK_HD LincombTerm lincomb_fetch(size_t i, uint32_t j, const LincombArgs& A) {
    LincombTerm t;
    const PolyDesc d = A.polys[j];
    t.kind = i < d.len ? d.kind : 0xFFFFFFFFu;
    t.x = fq_zero();
    t.y = fq_zero();
    if (t.kind == POLY_SHARED) {
        t.x = load_fq(d.data + 64 * i);
        t.y = load_fq(d.data + 64 * i + 32);
    } else if (t.kind != 0xFFFFFFFFu) {
        t.x = load_fq(d.data + 32 * i);
    }
    return t;
}
// thread i: element i of the joint polynomial.  The loads of term j+1 are issued before the multiplications of term j.
COZK_HD void lincomb_body(size_t i, const LincombArgs& A) {
    if (i >= A.n) return;
    fr_wide sa = fr_wide_zero(), sb = fr_wide_zero(), pub = fr_wide_zero();
    LincombTerm next = lincomb_fetch(i, 0, A);
    for (uint32_t j = 0; j < A.k; ++j) {
        const LincombTerm cur = next;
        if (j + 1 < A.k) next = lincomb_fetch(i, j + 1, A);
        if (cur.kind == POLY_SHARED) {
            fr c = load_fq(&A.coeffs[2 * j]);
            fr_wide_mac(sa, cur.x, c);
            fr_wide_mac(sb, cur.y, c);
        } else if (cur.kind != 0xFFFFFFFFu) {
            fr c = load_fq(&A.coeffs[2 * j + (cur.kind == POLY_CANON ? 1 : 0)]);
            fr_wide_mac(pub, cur.x, c);
        }
    }
    fr p = fr_wide_reduce(pub);
    if (!A.shared_out) {
        store_fq(A.out + 32 * i, p);
        return;
    }
    fr a = fr_wide_reduce(sa), b = fr_wide_reduce(sb);
    // rep3 add_public: the public value joins share a on party 0 and share b on party 1 (party 2 holds neither copy of t0)
    if (A.party == 0) a = fr_add(a, p);
    if (A.party == 1) b = fr_add(b, p);
    store_fq(A.out + 64 * i, a);
    store_fq(A.out + 64 * i + 32, b);
}

// chi dot product, stage 1: thread t of T sums its strided slice; out[poly * T + t] (Montgomery, reduced).
// Shared polynomial: (a_i + b_i) * chi_i (the 1/2 of into_additive is applied once, in stage 2); public: v_i * chi_i.
struct ChiArgs {
    const PolyDesc* polys;
    uint32_t k;
    const fr* chis;   // n values, Montgomery
    size_t n;
    uint32_t T;       // threads per polynomial
    fr* partial;      // k x T
};
COZK_HD fr chi_value(const PolyDesc& d, size_t i) {
    if (d.kind == POLY_SHARED) return fr_add(load_fq(d.data + 64 * i), load_fq(d.data + 64 * i + 32));
    fr v = load_fq(d.data + 32 * i);
    return d.kind == POLY_CANON ? fr_mont_from_canon(v) : v;
}
COZK_HD void chi_partial_body(size_t tid, const ChiArgs& A) {
    if (tid >= (size_t)A.k * A.T) return;
    uint32_t j = (uint32_t)(tid / A.T), t = (uint32_t)(tid - (size_t)j * A.T);
    const PolyDesc d = A.polys[j];
    fr_wide acc = fr_wide_zero();
    size_t lim = d.len < A.n ? d.len : A.n;
    size_t i = t;
    // four elements per round: their loads are all issued before the first multiplication
    for (; i + 3 * (size_t)A.T < lim; i += 4 * (size_t)A.T) {
        fr v0 = chi_value(d, i), v1 = chi_value(d, i + A.T), v2 = chi_value(d, i + 2 * (size_t)A.T), v3 = chi_value(d, i + 3 * (size_t)A.T);
        fr c0 = load_fq(&A.chis[i]), c1 = load_fq(&A.chis[i + A.T]), c2 = load_fq(&A.chis[i + 2 * (size_t)A.T]),
           c3 = load_fq(&A.chis[i + 3 * (size_t)A.T]);
        fr_wide_mac(acc, v0, c0);
        fr_wide_mac(acc, v1, c1);
        fr_wide_mac(acc, v2, c2);
        fr_wide_mac(acc, v3, c3);
    }
    for (; i < lim; i += A.T) fr_wide_mac(acc, chi_value(d, i), load_fq(&A.chis[i]));
    store_fq(&A.partial[tid], fr_wide_reduce(acc));
}
// stage 2 (run twice: T -> 64 -> 1 partial sums per polynomial, so that no thread walks a long chain of dependent loads):
// thread j * Tout + t adds the partial sums s = t, t + Tout, ... < Tin of polynomial j; the last pass applies the TWO_INV of
// into_additive for shared polynomials.
struct ChiReduceArgs {
    const PolyDesc* polys;
    uint32_t k;
    const fr* in;
    uint32_t Tin;
    fr* out;
    uint32_t Tout;
    uint32_t last;
};
COZK_HD void chi_reduce_body(size_t tid, const ChiReduceArgs& A) {
    if (tid >= (size_t)A.k * A.Tout) return;
    uint32_t j = (uint32_t)(tid / A.Tout), t = (uint32_t)(tid - (size_t)j * A.Tout);
    fr s = fq_zero();
    for (uint32_t i = t; i < A.Tin; i += A.Tout) s = fr_add(s, load_fq(&A.in[(size_t)j * A.Tin + i]));
    if (A.last && A.polys[j].kind == POLY_SHARED) {
        const uint32_t h[8] = COZK_FR_TWO_INV_MONT;
        s = fr_mul(s, fr_const(h));
    }
    store_fq(&A.out[tid], s);
}

// ------------------------------------------------------------------------------------------------ N1 pair sums
struct PairSumArgs {
    const affine* bases;      // 2 * half points
    const uint8_t* infinity;  // optional flags of the inputs
    size_t half;
    affine* out;              // half points
    uint8_t* out_inf;         // half flags (P + (-P), or both inputs at infinity)
};
COZK_HD void pair_sum_body(size_t b, const PairSumArgs& A) {
    if (b >= A.half) return;
    bool i0 = A.infinity && A.infinity[2 * b], i1 = A.infinity && A.infinity[2 * b + 1];
    xyzz s = xyzz_identity();
    if (!i0) s = xyzz_from_affine(load_affine(&A.bases[2 * b]));
    if (!i1) s = xyzz_madd(s, load_affine(&A.bases[2 * b + 1]));
    alignas(16) uint8_t w[72];
    xyzz_to_wire(s, w);
    affine r;
    const uint32_t* ww = reinterpret_cast<const uint32_t*>(w);
    for (int i = 0; i < 8; ++i) {
        r.x.v[i] = ww[i];
        r.y.v[i] = ww[8 + i];
    }
    store_fq(&A.out[b].x, r.x);
    store_fq(&A.out[b].y, r.y);
    A.out_inf[b] = w[64];
}

// ------------------------------------------------------------------------------------------------ partial joint polynomials
// A party's polynomials spread over several GPUs (one batch-commit shard per device): each device forms the linear
// combination of ITS polynomials, and the device that opens adds the partial results - the one exchange step of the
// commitment path.  Thread i reads element i of every partial (over NVLink peer mappings for the remote ones) and
// writes the joint element: transfer and addition are one kernel, nothing is staged.  Partials are lincomb outputs:
// shared (64 B) or public Montgomery (32 B); a public partial follows rep3 add_public like a public term.
struct SumPartialsArgs {
    const PolyDesc* parts;  // g descriptors; data may point into a peer device's memory
    uint32_t g;
    uint32_t party;
    uint32_t shared_out;
    size_t n;
    uint8_t* out;
};
COZK_HD void sum_partials_body(size_t i, const SumPartialsArgs& A) {
    if (i >= A.n) return;
    fr a = fq_zero(), b = fq_zero(), p = fq_zero();
    for (uint32_t j = 0; j < A.g; ++j) {
        const PolyDesc d = A.parts[j];
        if (i >= d.len) continue;
        if (d.kind == POLY_SHARED) {
            a = fr_add(a, load_fq(d.data + 64 * i));
            b = fr_add(b, load_fq(d.data + 64 * i + 32));
        } else {
            p = fr_add(p, load_fq(d.data + 32 * i));
        }
    }
    if (!A.shared_out) {
        store_fq(A.out + 32 * i, p);
        return;
    }
    if (A.party == 0) a = fr_add(a, p);
    if (A.party == 1) b = fr_add(b, p);
    store_fq(A.out + 64 * i, a);
    store_fq(A.out + 64 * i + 32, b);
}

// ------------------------------------------------------------------------------------------------ eq table
// chi[b] = prod_{i < nv} f_i(bit_i(b)),  f_i(1) = t_i, f_i(0) = 1 - t_i,  t_i = point[i] (LSB first) or point[nv-1-i]
// (MSB first).  Two small tables over the low lo_bits and the remaining high bits (one thread per entry, at most nv
// multiplications each), then ONE multiplication per element: chi[b] = hi[b >> lo_bits] * lo[b & mask].  All Montgomery.
struct EqArgs {
    const fr* point;   // nv values, Montgomery
    uint32_t nv, lo_bits;
    int msb_first;
    fr* lo_tab;        // [2^lo_bits]
    fr* hi_tab;        // [2^(nv - lo_bits)]
    fr* out;           // [2^nv]
};
COZK_HD fr eq_factor(const EqArgs& A, uint32_t i, uint32_t bit) {
    fr t = load_fq(&A.point[A.msb_first ? A.nv - 1 - i : i]);
    return bit ? t : fr_sub(fr_one(), t);
}
// threads [0, 2^lo_bits) fill lo_tab, the next 2^(nv - lo_bits) fill hi_tab
COZK_HD void eq_small_body(size_t tid, const EqArgs& A) {
    const size_t nlo = (size_t)1 << A.lo_bits, nhi = (size_t)1 << (A.nv - A.lo_bits);
    if (tid >= nlo + nhi) return;
    const bool hi = tid >= nlo;
    const size_t x = hi ? tid - nlo : tid;
    const uint32_t first = hi ? A.lo_bits : 0, count = hi ? A.nv - A.lo_bits : A.lo_bits;
    fr acc = fr_one();
    for (uint32_t i = 0; i < count; ++i) acc = fr_mul(acc, eq_factor(A, first + i, (uint32_t)(x >> i) & 1u));
    store_fq(hi ? &A.hi_tab[x] : &A.lo_tab[x], acc);
}
COZK_HD void eq_expand_body(size_t b, const EqArgs& A) {
    if (b >= ((size_t)1 << A.nv)) return;
    const size_t mask = ((size_t)1 << A.lo_bits) - 1;
    store_fq(&A.out[b], fr_mul(load_fq(&A.hi_tab[b >> A.lo_bits]), load_fq(&A.lo_tab[b & mask])));
}

}  // namespace cozk
