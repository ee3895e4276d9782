// Batched-affine pre-reduction of the sorted (bucket key, point) pair list, in front of the XYZZ accumulate levels.
//
// A mixed XYZZ addition costs 8M + 2S (1232 limb products as issued here).  Two AFFINE points add in 2M + 1S plus one
// inversion, and Montgomery's trick shares one inversion among any number of independent additions for 3M each:
// 5M + 1S = 788 limb products per addition.  Buckets hold dozens to hundreds of points (2^20 points, window 17: 240 per
// bucket), so neighbours in the sorted list are independent additions of the same bucket:
//
//   round r:  out[o] = in[2o] + in[2o + 1]   when both carry the same key        (half the list, still grouped by key)
//             a pair (2o, 2o + 1) that straddles two buckets cannot be added: out[o] = in[2o], and in[2o + 1] - the first
//             point of its bucket's run - goes to the round's OVERFLOW list (at most one entry per bucket and round), which
//             is added to the buckets after the accumulate levels (k_ovf_add: one mixed addition per entry).
//
// After R rounds (R = 3: 7/8 of all additions done at 788 instead of 1232) the list - now (key, index into the round's
// output points) - goes through the unchanged XYZZ accumulate levels.  Every exceptional case is exact: P + P is a
// doubling with its own denominator 2y, P + (-P) and skipped pairs (zero digits, bases at infinity) are null entries.
//
// One round on the GPU is three launches over batches of AFF_T * AFF_K additions:
//   k_affine_prod   thread t of a batch multiplies the denominators of its AFF_K additions; the block multiplies the thread
//                   products: ONE product per batch
//   k_affine_inv    inverts the batch products, one lane per batch (Fermat, every lane busy)
//   k_affine_apply  the same forward pass again, this time keeping the prefix products in shared memory; the thread
//                   products' inverses from the batch inverse by a product tree (warp 0: 4 per lane, butterfly over the
//                   lanes); then the backward pass: inverse of every denominator, slope, sum.
// Output contract of the three = affine_round_body (one addition with its own inversion), which is what the host
// emulation (tests/emul) and the GPU test kernel k_affine_ref run.  The order of the overflow list is unspecified.
#pragma once
#include "msm_kernels.cuh"

namespace cozk {

constexpr uint32_t AFF_NOP = 0, AFF_PASS = 1, AFF_ADD = 2, AFF_DBL = 3;

struct AffineRoundArgs {
    size_t m;                  // input entries
    const uint32_t* keys_in;   // [m] grouped by key
    const uint32_t* vals_in;   // [m] index into pts_in | sign << 31, or VAL_SKIP
    const affine* pts_in;
    uint32_t* keys_out;        // [(m + 1) / 2]
    uint32_t* vals_out;        // [(m + 1) / 2]: o, or VAL_SKIP for a null entry
    affine* pts_out;           // [(m + 1) / 2]
    uint32_t* ovf_count;       // entries of this round's overflow list (may exceed ovf_cap: the list is then incomplete)
    uint32_t* ovf_keys;        // [ovf_cap]
    affine* ovf_pts;           // [ovf_cap]
    uint32_t ovf_cap;
    fq* batch_prod;            // [batches] (batched form only)
    fq* batch_inv;             // [batches]
};
COZK_HD size_t affine_round_out(size_t m) { return (m + 1) / 2; }

COZK_HD bool affine_key_null(uint32_t key) { return (key & KEY_MASK) == KEY_MASK; }
COZK_HD affine affine_load_signed(const affine* pts, uint32_t val) {
    affine p = load_affine(&pts[val & ~VAL_NEG]);
    p.y = fq_cneg(p.y, (val & VAL_NEG) != 0);
    return p;
}

// What output o has to do, and for a real addition its denominator d (AFF_ADD: x1 - x0, AFF_DBL: 2 y0; never zero).
COZK_HD uint32_t affine_classify(size_t o, const AffineRoundArgs& A, fq& d) {
    if (o >= affine_round_out(A.m)) return AFF_NOP;
    const size_t i0 = 2 * o;
    if (i0 + 1 >= A.m) return AFF_PASS;
    const uint32_t k0 = A.keys_in[i0], k1 = A.keys_in[i0 + 1];
    const uint32_t v0 = A.vals_in[i0], v1 = A.vals_in[i0 + 1];
    if (k0 != k1 || v0 == VAL_SKIP || v1 == VAL_SKIP || affine_key_null(k0)) return AFF_PASS;
    const fq x0 = load_fq(&A.pts_in[v0 & ~VAL_NEG].x), x1 = load_fq(&A.pts_in[v1 & ~VAL_NEG].x);
    if (!fq_eq(x0, x1)) {
        d = fq_sub(x1, x0);
        return AFF_ADD;
    }
    const fq y0 = fq_cneg(load_fq(&A.pts_in[v0 & ~VAL_NEG].y), (v0 & VAL_NEG) != 0);
    const fq y1 = fq_cneg(load_fq(&A.pts_in[v1 & ~VAL_NEG].y), (v1 & VAL_NEG) != 0);
    if (fq_eq(y0, y1) && !fq_is_zero(y0)) {
        d = fq_dbl(y0);
        return AFF_DBL;
    }
    return AFF_PASS;  // P + (-P): the identity, a null entry
}

// the sum, given 1 / d
COZK_HD affine affine_finish(uint32_t code, const affine& p0, const affine& p1, const fq& inv_d) {
    fq num;
    if (code == AFF_ADD) {
        num = fq_sub(p1.y, p0.y);
    } else {
        const fq xx = fq_sqr(p0.x);
        num = fq_add(fq_dbl(xx), xx);
    }
    const fq lam = fq_mul(num, inv_d);
    affine r;
    r.x = fq_sub(fq_sub(fq_sqr(lam), p0.x), p1.x);
    r.y = fq_sub(fq_mul(lam, fq_sub(p0.x, r.x)), p0.y);
    return r;
}

COZK_HD void affine_push_overflow(const AffineRoundArgs& A, uint32_t key, const affine& p) {
#if defined(__CUDA_ARCH__)
    const uint32_t s = atomicAdd(A.ovf_count, 1u);
#else
    const uint32_t s = (*A.ovf_count)++;
#endif
    if (s < A.ovf_cap) {
        A.ovf_keys[s] = key;
        store_fq(&A.ovf_pts[s].x, p.x);
        store_fq(&A.ovf_pts[s].y, p.y);
    }
}
COZK_HD void affine_store_out(size_t o, const AffineRoundArgs& A, uint32_t key, bool null, const affine& p) {
    A.keys_out[o] = key;
    A.vals_out[o] = null ? VAL_SKIP : (uint32_t)o;
    if (!null) {
        store_fq(&A.pts_out[o].x, p.x);
        store_fq(&A.pts_out[o].y, p.y);
    }
}

// an output without an addition: a lone last entry, a pair across two buckets, null operands, P + (-P)
COZK_HD void affine_pass(size_t o, const AffineRoundArgs& A) {
    const size_t i0 = 2 * o;
    const bool has1 = i0 + 1 < A.m;
    const uint32_t k0 = A.keys_in[i0], v0 = A.vals_in[i0];
    const uint32_t k1 = has1 ? A.keys_in[i0 + 1] : k0, v1 = has1 ? A.vals_in[i0 + 1] : VAL_SKIP;
    const bool n0 = v0 == VAL_SKIP || affine_key_null(k0), n1 = v1 == VAL_SKIP || affine_key_null(k1);
    affine p;
    p.x = fq_zero();
    p.y = fq_zero();
    bool null = true;
    if (k0 != k1) {
        if (!n1) affine_push_overflow(A, k1, affine_load_signed(A.pts_in, v1));
        if (!n0) {
            p = affine_load_signed(A.pts_in, v0);
            null = false;
        }
    } else if (n0 != n1) {
        p = affine_load_signed(A.pts_in, n0 ? v1 : v0);
        null = false;
    }
    affine_store_out(o, A, k0, null, p);
}

// Output contract of one round, thread o: its own inversion per addition (host emulation, GPU test reference).
COZK_HD void affine_round_body(size_t o, const AffineRoundArgs& A) {
    fq d;
    const uint32_t code = affine_classify(o, A, d);
    if (code == AFF_NOP) return;
    if (code == AFF_PASS) {
        affine_pass(o, A);
        return;
    }
    const uint32_t v0 = A.vals_in[2 * o], v1 = A.vals_in[2 * o + 1];
    const affine p0 = affine_load_signed(A.pts_in, v0), p1 = affine_load_signed(A.pts_in, v1);
    affine_store_out(o, A, A.keys_in[2 * o], false, affine_finish(code, p0, p1, fq_inv(d)));
}

// The overflow list of a round joins the buckets behind the accumulate levels: keys are distinct inside one list.
struct OvfAddArgs {
    const uint32_t* count;
    const uint32_t* keys;
    const affine* pts;
    xyzz* buckets;
    uint32_t cap;
};
COZK_HD void ovf_add_body(size_t i, const OvfAddArgs& A) {
    const uint32_t n = *A.count < A.cap ? *A.count : A.cap;
    if (i >= n) return;
    const uint32_t key = A.keys[i];
    if (affine_key_null(key)) return;
    store_xyzz(&A.buckets[key], xyzz_madd(load_xyzz(&A.buckets[key]), load_affine(&A.pts[i])));
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------ batched kernels
#ifndef COZK_AFF_T
#define COZK_AFF_T 128
#endif
#ifndef COZK_AFF_K
#define COZK_AFF_K 16
#endif
#ifndef COZK_AFF_PREFETCH
#define COZK_AFF_PREFETCH 0
#endif
#ifndef COZK_AFF_MINBLOCKS
#define COZK_AFF_MINBLOCKS 3
#endif
constexpr int AFF_T = COZK_AFF_T;   // threads per batch
constexpr int AFF_K = COZK_AFF_K;   // additions per thread
constexpr int AFF_BATCH = AFF_T * AFF_K;

struct AffineSmem {
    uint4 pre[(AFF_K - 1) * 2 * AFF_T];  // prefix product k (of denominators 0 .. k) of thread t: pre[(k * 2 + half) * AFF_T + t]
    uint4 red[2 * AFF_T];                // thread products, then their inverses: red[half * AFF_T + t]
};

__device__ __forceinline__ void aff_st(uint4* base, int t, const fq& a) {
    base[t] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    base[AFF_T + t] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
__device__ __forceinline__ fq aff_ld(const uint4* base, int t) {
    const uint4 a = base[t], b = base[AFF_T + t];
    fq r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ fq aff_shfl_xor(const fq& a, int s) {
    fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __shfl_xor_sync(0xFFFFFFFFu, a.v[i], s);
    return r;
}

// output handled by thread t at step k of batch `batch`: consecutive threads take consecutive outputs
__device__ __forceinline__ size_t aff_out_index(size_t batch, int k, int t) { return batch * AFF_BATCH + (size_t)k * AFF_T + t; }

// What the forward pass needs of output o, loaded without looking at it: the loads of several steps are in flight together.
struct AffLoaded {
    uint32_t state;  // AFF_NOP, AFF_PASS, or AFF_ADD = a pair of live entries of one bucket whose x coordinates are loaded
    uint32_t v0, v1;
    fq x0, x1;
};
__device__ __forceinline__ void aff_issue(size_t o, const AffineRoundArgs& A, AffLoaded& L) {
    L.state = AFF_NOP;
    if (o >= affine_round_out(A.m)) return;
    L.state = AFF_PASS;
    const size_t i0 = 2 * o;
    if (i0 + 1 >= A.m) return;
    const uint2 kk = *reinterpret_cast<const uint2*>(A.keys_in + i0), vv = *reinterpret_cast<const uint2*>(A.vals_in + i0);
    L.v0 = vv.x;
    L.v1 = vv.y;
    if (kk.x != kk.y || vv.x == VAL_SKIP || vv.y == VAL_SKIP || affine_key_null(kk.x)) return;
    L.state = AFF_ADD;
    L.x0 = load_fq(&A.pts_in[vv.x & ~VAL_NEG].x);
    L.x1 = load_fq(&A.pts_in[vv.y & ~VAL_NEG].x);
}
// same decision and denominator as affine_classify
__device__ __forceinline__ uint32_t aff_decide(const AffLoaded& L, const AffineRoundArgs& A, fq& d) {
    if (L.state != AFF_ADD) return L.state;
    if (!fq_eq(L.x0, L.x1)) {
        d = fq_sub(L.x1, L.x0);
        return AFF_ADD;
    }
    const fq y0 = fq_cneg(load_fq(&A.pts_in[L.v0 & ~VAL_NEG].y), (L.v0 & VAL_NEG) != 0);
    const fq y1 = fq_cneg(load_fq(&A.pts_in[L.v1 & ~VAL_NEG].y), (L.v1 & VAL_NEG) != 0);
    if (fq_eq(y0, y1) && !fq_is_zero(y0)) {
        d = fq_dbl(y0);
        return AFF_DBL;
    }
    return AFF_PASS;
}

// Forward pass of one thread: the product of its AFF_K denominators (1 for an output without an addition); KEEP: the prefix
// products go to shared memory.  codes: 2 bits per step.
template <bool KEEP>
__device__ __forceinline__ fq aff_forward(const AffineRoundArgs& A, size_t batch, int t, AffineSmem* S, uint32_t& codes) {
    fq run = fq_one();
    codes = 0;
#if COZK_AFF_PREFETCH
    constexpr int G = 4;  // steps whose loads are in flight together
    static_assert(AFF_K % G == 0, "AFF_K must be a multiple of 4");
#pragma unroll 1
    for (int k0 = 0; k0 < AFF_K; k0 += G) {
        AffLoaded L[G];
#pragma unroll
        for (int j = 0; j < G; ++j) aff_issue(aff_out_index(batch, k0 + j, t), A, L[j]);
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int k = k0 + j;
            fq d;
            const uint32_t code = aff_decide(L[j], A, d);
            codes |= code << (2 * k);
            if (KEEP && k > 0) aff_st(S->pre + (size_t)(k - 1) * 2 * AFF_T, t, run);
            if (code >= AFF_ADD) run = k == 0 ? d : fq_mul(run, d);
        }
    }
#else
#pragma unroll 1
    for (int k = 0; k < AFF_K; ++k) {
        fq d;
        const uint32_t code = affine_classify(aff_out_index(batch, k, t), A, d);
        codes |= code << (2 * k);
        if (KEEP && k > 0) aff_st(S->pre + (size_t)(k - 1) * 2 * AFF_T, t, run);
        if (code >= AFF_ADD) run = k == 0 ? d : fq_mul(run, d);
    }
#endif
    return run;
}

// product of the AFF_T thread products (in red), by warp 0 (every lane ends with it).  Block-synchronised on entry.
__device__ __forceinline__ fq aff_block_product(const uint4* red) {
    const int lane = threadIdx.x & 31;
    constexpr int PER = AFF_T / 32;
    fq v = aff_ld(red, lane * PER);
#pragma unroll
    for (int j = 1; j < PER; ++j) v = fq_mul(v, aff_ld(red, lane * PER + j));
#pragma unroll 1
    for (int s = 1; s < 32; s <<= 1) v = fq_mul(v, aff_shfl_xor(v, s));
    return v;
}

__global__ void __launch_bounds__(AFF_T, COZK_AFF_MINBLOCKS) k_affine_prod(AffineRoundArgs A) {
    __shared__ uint4 red[2 * AFF_T];
    const int t = threadIdx.x;
    uint32_t codes;
    const fq p = aff_forward<false>(A, blockIdx.x, t, nullptr, codes);
    aff_st(red, t, p);
    __syncthreads();
    if (t < 32) {
        const fq v = aff_block_product(red);
        if (t == 0) store_fq(&A.batch_prod[blockIdx.x], v);
    }
}

__global__ void __launch_bounds__(128) k_affine_inv(const fq* prod, fq* inv, uint32_t batches) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batches) return;
    store_fq(&inv[b], fq_inv(load_fq(&prod[b])));
}

__global__ void __launch_bounds__(AFF_T, COZK_AFF_MINBLOCKS) k_affine_apply(AffineRoundArgs A) {
    extern __shared__ __align__(16) unsigned char aff_smem_raw[];
    AffineSmem* S = reinterpret_cast<AffineSmem*>(aff_smem_raw);
    const int t = threadIdx.x;
    const size_t batch = blockIdx.x;
    uint32_t codes;
    const fq p = aff_forward<true>(A, batch, t, S, codes);
    aff_st(S->red, t, p);
    __syncthreads();
    if (t < 32) {
        // inverse of every thread product from the inverse of the batch product: lane l owns PER products, the lanes'
        // totals meet in a butterfly that also gathers, per lane, the product of all the OTHER lanes
        constexpr int PER = AFF_T / 32;
        fq a[PER], pp[PER];  // pp[j] = a[0] .. a[j]
#pragma unroll
        for (int j = 0; j < PER; ++j) a[j] = aff_ld(S->red, t * PER + j);
        pp[0] = a[0];
#pragma unroll
        for (int j = 1; j < PER; ++j) pp[j] = fq_mul(pp[j - 1], a[j]);
        fq v = pp[PER - 1], oth = fq_one();
#pragma unroll 1
        for (int s = 1; s < 32; s <<= 1) {
            const fq q = aff_shfl_xor(v, s);
            oth = s == 1 ? q : fq_mul(oth, q);
            v = fq_mul(v, q);
        }
        fq inv = fq_mul(load_fq(&A.batch_inv[batch]), oth);  // 1 / pp[PER - 1]
#pragma unroll
        for (int j = PER - 1; j >= 1; --j) {
            aff_st(S->red, t * PER + j, fq_mul(inv, pp[j - 1]));  // 1 / a[j]
            inv = fq_mul(inv, a[j]);                              // 1 / pp[j - 1]
        }
        aff_st(S->red, t * PER, inv);
    }
    __syncthreads();
    fq inv_run = aff_ld(S->red, t);  // 1 / (d_0 .. d_{K-1}) of this thread
#if COZK_AFF_PREFETCH
    // the points of step k - 1 are requested before the arithmetic of step k starts (raw: the sign is applied where they are used)
    affine q0, q1;
    uint32_t qv0 = 0, qv1 = 0;
    q0.x = q0.y = q1.x = q1.y = fq_zero();
    if (((codes >> (2 * (AFF_K - 1))) & 3u) >= AFF_ADD) {
        const size_t o = aff_out_index(batch, AFF_K - 1, t);
        const uint2 vv = *reinterpret_cast<const uint2*>(A.vals_in + 2 * o);
        qv0 = vv.x, qv1 = vv.y;
        q0 = load_affine(&A.pts_in[qv0 & ~VAL_NEG]);
        q1 = load_affine(&A.pts_in[qv1 & ~VAL_NEG]);
    }
#endif
#pragma unroll 1
    for (int k = AFF_K - 1; k >= 0; --k) {
        const uint32_t code = (codes >> (2 * k)) & 3u;
        const size_t o = aff_out_index(batch, k, t);
#if COZK_AFF_PREFETCH
        affine p0 = q0, p1 = q1;
        const uint32_t pv0 = qv0, pv1 = qv1;
        if (k > 0 && ((codes >> (2 * (k - 1))) & 3u) >= AFF_ADD) {
            const size_t on = aff_out_index(batch, k - 1, t);
            const uint2 vv = *reinterpret_cast<const uint2*>(A.vals_in + 2 * on);
            qv0 = vv.x, qv1 = vv.y;
            q0 = load_affine(&A.pts_in[qv0 & ~VAL_NEG]);
            q1 = load_affine(&A.pts_in[qv1 & ~VAL_NEG]);
        }
        if (code >= AFF_ADD) {
            p0.y = fq_cneg(p0.y, (pv0 & VAL_NEG) != 0);
            p1.y = fq_cneg(p1.y, (pv1 & VAL_NEG) != 0);
        }
#endif
        if (code >= AFF_ADD) {
#if !COZK_AFF_PREFETCH
            const uint32_t v0 = A.vals_in[2 * o], v1 = A.vals_in[2 * o + 1];
            const affine p0 = affine_load_signed(A.pts_in, v0), p1 = affine_load_signed(A.pts_in, v1);
#endif
            const fq d = code == AFF_ADD ? fq_sub(p1.x, p0.x) : fq_dbl(p0.y);
            fq inv_d = inv_run;
            if (k > 0) {
                inv_d = fq_mul(inv_run, aff_ld(S->pre + (size_t)(k - 1) * 2 * AFF_T, t));
                inv_run = fq_mul(inv_run, d);
            }
            affine_store_out(o, A, A.keys_in[2 * o], false, affine_finish(code, p0, p1, inv_d));
        } else if (code == AFF_PASS) {
            affine_pass(o, A);
        }
    }
}

__global__ void __launch_bounds__(128) k_affine_ref(AffineRoundArgs A) { affine_round_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
#endif  // __CUDACC__

}  // namespace cozk
