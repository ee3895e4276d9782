// libcozk_test.so - TEST AND MEASUREMENT entry points (include/cozk_test.h): on-device synthetic input generation
// (SURVEY.md section 8(d)), element-wise kernels for parity tests, the pair sort on its own, and the roofline
// microbenchmarks (self-measured integer-multiply pipe peak; MEASURED_PEAKS.json has no integer figure).  Links against
// libcozk_msm.so for the engine's internals; nothing in the product library depends on this file.
#include "../../include/cozk_test.h"
#include "engine.hpp"
#include "affine_kernels.cuh"
#include "msm_kernels.cuh"

namespace cozk {

// ------------------------------------------------------------------------------------------------ generators
__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t gen_limb(uint64_t seed, uint64_t i, uint64_t j) { return mix64(mix64(seed) + 4 * i + j); }

// 254-bit value from the stream, reduced once below `mod` (mod > 2^253)
__device__ inline fq raw254(uint64_t seed, uint64_t i, const uint32_t* mod) {
    fq r;
    for (int j = 0; j < 4; ++j) {
        uint64_t l = gen_limb(seed, i, j);
        if (j == 3) l &= 0x3FFFFFFFFFFFFFFFULL;
        r.v[2 * j] = (uint32_t)l;
        r.v[2 * j + 1] = (uint32_t)(l >> 32);
    }
    uint32_t s[8];
    uint32_t br = 0;
    for (int k = 0; k < 8; ++k) {
        uint64_t d = (uint64_t)r.v[k] - mod[k] - br;
        s[k] = (uint32_t)d;
        br = (uint32_t)(d >> 32) & 1u;
    }
    if (!br)
        for (int k = 0; k < 8; ++k) r.v[k] = s[k];
    return r;
}

__device__ inline fq fq_to_mont(const fq& a) {
    const uint32_t r2[8] = COZK_FQ_R2;
    fq t;
    for (int i = 0; i < 8; ++i) t.v[i] = r2[i];
    return fq_mul(a, t);
}
__device__ inline fq fq_from_mont(const fq& a) {
    fq one = fq_zero();
    one.v[0] = 1;
    return fq_mul(a, one);
}
// a^((p+1)/4)
__device__ inline fq fq_sqrt_candidate(const fq& a) {
    const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
    fq acc = fq_one();
    for (int i = 251; i >= 0; --i) {
        acc = fq_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fq_mul(acc, a);
    }
    return acc;
}

__global__ void k_gen_bases(uint64_t seed, size_t start, size_t n, affine* out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t mod[8] = COZK_FQ_MOD;
    fq x = fq_to_mont(raw254(seed, start + t, mod));
    fq one = fq_one();
    fq three = fq_add(fq_add(one, one), one);
    fq y;
    for (int tries = 0; tries < 256; ++tries) {  // half of all x work; the bound only guards against a hung kernel
        fq rhs = fq_add(fq_mul(fq_sqr(x), x), three);
        y = fq_sqrt_candidate(rhs);
        if (fq_eq(fq_sqr(y), rhs)) break;
        x = fq_add(x, one);
    }
    if (fq_from_mont(y).v[0] & 1u) y = fq_neg(y);
    store_fq(&out[t].x, x);
    store_fq(&out[t].y, y);
}

// Fr helpers for the scalar distributions
__device__ inline fr fr_sub_mod(const fr& a, const fr& b) {
    const uint32_t mod[8] = COZK_FR_MOD;
    fr t;
    uint32_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        t.v[i] = (uint32_t)d;
        br = (uint32_t)(d >> 32) & 1u;
    }
    if (br) {
        uint32_t c = 0;
        for (int i = 0; i < 8; ++i) {
            uint64_t s = (uint64_t)t.v[i] + mod[i] + c;
            t.v[i] = (uint32_t)s;
            c = (uint32_t)(s >> 32);
        }
    }
    return t;
}
__device__ inline fr fr_to_mont(const fr& a) {
    // R^2 mod r
    const uint32_t r2[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    fr t;
    for (int i = 0; i < 8; ++i) t.v[i] = r2[i];
    return fr_mul(a, t);
}

__global__ void k_gen_scalars(int dist, uint64_t seed, size_t start, size_t n, size_t total_n, int form, uint8_t* out, size_t stride) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t mod[8] = COZK_FR_MOD;
    size_t i = start + t;
    fr s = fq_zero();
    switch (dist) {
        case 0: s = raw254(seed, i, mod); break;
        case 1: s = raw254(seed, 0, mod); break;
        case 2: {
            fr w = fq_zero();
            w.v[0] = (uint32_t)gen_limb(seed, i + 2, 0);
            s = fr_sub_mod(fr_sub_mod(w, raw254(seed, 0, mod)), raw254(seed, 1, mod));
            break;
        }
        case 3: s = raw254(seed, i >> 1, mod); break;
        case 4: s.v[0] = (uint32_t)gen_limb(seed, i, 0) & 0xFFFFu; break;
        default: if (i < (total_n + 1) / 2) s = raw254(seed, i, mod); break;
    }
    if (form == COZK_MONT) s = fr_to_mont(s);
    store_fq(out + t * stride, s);
}

// ------------------------------------------------------------------------------------------------ element-wise test kernels
__global__ void k_field_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fq x = load_fq(a + 32 * t), y = b ? load_fq(b + 32 * t) : fq_zero(), r;
    switch (op) {
        case 0: r = fq_mul(x, y); break;
        case 1: r = fq_add(x, y); break;
        case 2: r = fq_sub(x, y); break;
        case 3: r = fq_sqr(x); break;
        case 4: r = fq_inv(x); break;
        case 7: r = fq_mul2(x, y, y, x); break;  // 2xy with one reduction
        default: r = fr_from_mont(x); break;
    }
    store_fq(out + 32 * t, r);
}

__global__ void k_g1_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    xyzz pa = xyzz_from_wire(a + 72 * t), r;
    if (op == 0) {
        r = xyzz_add(pa, xyzz_from_wire(b + 72 * t));
    } else if (op == 1) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(b + 72 * t);
        affine q;
        for (int i = 0; i < 8; ++i) {
            q.x.v[i] = w[i];
            q.y.v[i] = w[8 + i];
        }
        r = xyzz_madd(pa, q);
    } else {
        r = xyzz_dbl(pa);
    }
    xyzz_to_wire(r, out + 72 * t);
}

// ------------------------------------------------------------------------------------------------ microbenchmarks
// 0: eight chains of 32 x 32 + 64 -> 64 multiply-accumulates per thread.  Each multiplicand is the low word of ANOTHER
//    chain's accumulator, so no product is loop-invariant (ptxas otherwise hoists the multiply and the loop measures
//    IADD3).  ptxas lowers mad.wide.u32 with a live addend to IMAD.WIDE.U32 R,a,b,RZ + a 3-input IADD3 / IADD3.X pair.
__global__ void k_bench_imad(int iters, uint32_t* sink) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint64_t acc[8];
    for (int k = 0; k < 8; ++k) acc[k] = (uint64_t)k * 0x9E3779B97F4A7C15ULL + a + b;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t mul = (uint32_t)acc[(k + 3) & 7];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(mul), "r"(b));
        }
    }
    uint64_t s = 0;
    for (int k = 0; k < 8; ++k) s ^= acc[k];
    if (s == 0x1234567ULL) sink[0] = (uint32_t)s;
}
// 1 / 2: dependent chains of field multiplications / squarings, two independent chains per thread
__global__ void k_bench_fq(int which, int iters, const uint8_t* in, uint8_t* sink) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fq x = load_fq(in + 32 * (t & 1023)), y = load_fq(in + 32 * ((t + 7) & 1023));
    fq u = y, v = x;
    if (which == 1) {
        for (int i = 0; i < iters; ++i) {
            x = fq_mul(x, y);
            u = fq_mul(u, v);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
            x = fq_sqr(x);
            u = fq_sqr(u);
        }
    }
    fq r = fq_add(x, u);
    if (r.v[0] == 0x12345u && r.v[7] == 0x7777u) store_fq(sink, r);
}
// 3: a chain of mixed additions from registers (no memory traffic): the accumulate kernel's arithmetic ceiling
__global__ void __launch_bounds__(128, 4) k_bench_madd(int iters, const uint8_t* in, uint8_t* sink) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    affine p;
    p.x = load_fq(in + 64 * (t & 511));
    p.y = load_fq(in + 64 * (t & 511) + 32);
    affine q;
    q.x = load_fq(in + 64 * ((t + 1) & 511));
    q.y = load_fq(in + 64 * ((t + 1) & 511) + 32);
    xyzz acc = xyzz_from_affine(p);
    for (int i = 0; i < iters; ++i) {
        acc = xyzz_madd(acc, q);
        q.y = fq_neg(q.y);
        acc = xyzz_madd(acc, p);
    }
    if (acc.X.v[0] == 0x12345u && acc.ZZ.v[7] == 0x7777u) store_xyzz(reinterpret_cast<xyzz*>(sink), acc);
}

// 4: the carry-chain form the field multiplication actually uses: rows of (mad.lo.cc, madc.hi.cc) pairs, which ptxas
//    fuses into IMAD.WIDE.U32.X with a carry predicate in and out.  4 chains of 4 pairs per iteration; the multiplier
//    of each chain is a word of a neighbouring chain, so nothing is loop-invariant.  This is the roofline denominator:
//    the highest rate at which the chip retires 32 x 32 -> 64-bit multiply-accumulates.
__global__ void k_bench_imad_cc(int iters, uint32_t* sink) {
    uint32_t a0 = threadIdx.x * 2654435761u + 12345u, a1 = a0 ^ 0x55aa55aau, a2 = a0 * 3u + 1u, a3 = a0 * 7u + 5u;
    uint32_t e[4][9];
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 9; ++k) e[c][k] = a0 + 17u * k + c + blockIdx.x * 40503u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            asm volatile(
                "mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
                "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
                "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
                "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
                "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
                "addc.u32 %8, %8, 0;\n\t"
                : "+r"(e[c][0]), "+r"(e[c][1]), "+r"(e[c][2]), "+r"(e[c][3]), "+r"(e[c][4]), "+r"(e[c][5]), "+r"(e[c][6]),
                  "+r"(e[c][7]), "+r"(e[c][8])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(e[(c + 1) & 3][7] | 1u));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 9; ++k) s ^= e[c][k];
    if (s == 0x1234567u) sink[0] = s;
}
// 5 / 6: 32-bit mad.lo / mad.hi, eight independent chains
__global__ void k_bench_imad32(int hi, int iters, uint32_t* sink) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint32_t acc[8];
    for (int k = 0; k < 8; ++k) acc[k] = a * (k + 3);
    if (hi) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 3) & 7] | 0x80000001u), "r"(b));
        }
    } else {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 3) & 7] | 1u), "r"(b));
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < 8; ++k) s ^= acc[k];
    if (s == 0x1234567u) sink[0] = s;
}
// 7: four independent field-multiplication chains per thread (more ILP than bench 1)
__global__ void k_bench_fq4(int iters, const uint8_t* in, uint8_t* sink) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fq x0 = load_fq(in + 32 * (t & 1023)), x1 = load_fq(in + 32 * ((t + 7) & 1023));
    fq x2 = load_fq(in + 32 * ((t + 13) & 1023)), x3 = load_fq(in + 32 * ((t + 29) & 1023));
    fq y = load_fq(in + 32 * ((t + 3) & 1023));
    for (int i = 0; i < iters; ++i) {
        x0 = fq_mul(x0, y);
        x1 = fq_mul(x1, y);
        x2 = fq_mul(x2, y);
        x3 = fq_mul(x3, y);
    }
    fq r = fq_add(fq_add(x0, x1), fq_add(x2, x3));
    if (r.v[0] == 0x12345u && r.v[7] == 0x7777u) store_fq(sink, r);
}

static int get_device(cozk_ctx* ctx, int device_index, Device** out) {
    if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size()) {
        set_error("bad context or device index");
        return COZK_ERR_INVALID_ARG;
    }
    *out = ctx->devs[device_index].get();
    COZK_CUDA(cudaSetDevice((*out)->id));
    return COZK_OK;
}

}  // namespace cozk

using namespace cozk;

extern "C" {

int cozk_testgen_bases(cozk_ctx* ctx, int device_index, uint64_t seed, size_t start, size_t n, void* d_out64) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    if (n == 0) return COZK_OK;
    k_gen_bases<<<(unsigned)((n + 127) / 128), 128, 0, D->stream>>>(seed, start, n, reinterpret_cast<affine*>(d_out64));
    COZK_CUDA(cudaGetLastError());
    COZK_CUDA(cudaStreamSynchronize(D->stream));
    return COZK_OK;
}
int cozk_testgen_scalars(cozk_ctx* ctx, int device_index, int dist, uint64_t seed, size_t start, size_t n, size_t total_n,
                         int form, void* d_out, size_t stride_bytes) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    if (n == 0) return COZK_OK;
    if (stride_bytes < 32 || (stride_bytes & 15) || dist < 0 || dist > 5) {
        set_error("bad stride or distribution");
        return COZK_ERR_INVALID_ARG;
    }
    k_gen_scalars<<<(unsigned)((n + 255) / 256), 256, 0, D->stream>>>(dist, seed, start, n, total_n, form,
                                                                      reinterpret_cast<uint8_t*>(d_out), stride_bytes);
    COZK_CUDA(cudaGetLastError());
    COZK_CUDA(cudaStreamSynchronize(D->stream));
    return COZK_OK;
}

int cozk_test_field_op(cozk_ctx* ctx, int device_index, int op, const void* d_a, const void* d_b, void* d_out, size_t n) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    if (n == 0) return COZK_OK;
    k_field_op<<<(unsigned)((n + 127) / 128), 128, 0, D->stream>>>(op, (const uint8_t*)d_a, (const uint8_t*)d_b, (uint8_t*)d_out, n);
    COZK_CUDA(cudaGetLastError());
    COZK_CUDA(cudaStreamSynchronize(D->stream));
    return COZK_OK;
}
int cozk_test_g1_op(cozk_ctx* ctx, int device_index, int op, const void* d_a, const void* d_b, void* d_out, size_t n) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    if (n == 0) return COZK_OK;
    k_g1_op<<<(unsigned)((n + 63) / 64), 64, 0, D->stream>>>(op, (const uint8_t*)d_a, (const uint8_t*)d_b, (uint8_t*)d_out, n);
    COZK_CUDA(cudaGetLastError());
    COZK_CUDA(cudaStreamSynchronize(D->stream));
    return COZK_OK;
}

// The pair sort on its own.  d_scalars == NULL: groups the m given (key, val) pairs by key, keys ascending (keys below
// 2^key_bits; order inside a group unspecified).  d_scalars != NULL: the pairs are those of the plain decompose layout of g vectors of n scalars (window c,
// `windows` windows, table_stride / val_offset as in DecomposeArgs); fused != 0 produces them inside the first sort pass
// (the engine's path), fused == 0 with the decompose kernel followed by generic passes; key_bits == 0 leaves them unsorted
// (fused == 0 only).  Outputs: m = g * n * windows pairs.
int cozk_test_sort(cozk_ctx* ctx, int device_index, const void* d_keys, const void* d_vals, size_t m, unsigned key_bits,
                   const void* d_scalars, size_t n, unsigned g, size_t stride, int form, unsigned c, unsigned windows,
                   size_t table_stride, size_t val_offset, int fused, void* d_keys_out, void* d_vals_out) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(D->mu);
    D->sort_digit_bits = ctx->opt_sort_digit_bits;
    if (d_scalars) m = (size_t)g * n * windows;
    if (m == 0) return COZK_OK;
    if ((rc = D->keys_a.ensure(m * 4)) || (rc = D->vals_a.ensure(m * 4)) || (rc = D->keys_b.ensure(m * 4)) || (rc = D->vals_b.ensure(m * 4)))
        return rc;
    cudaStream_t st = D->stream;
    uint32_t *ks = D->keys_a.as<uint32_t>(), *vs = D->vals_a.as<uint32_t>();
    double launches = 0;
    if (d_scalars) {
        const size_t vstride = ((n - 1) * stride + 32 + 255) & ~(size_t)255;
        DecomposeArgs DA{reinterpret_cast<const uint8_t*>(d_scalars), nullptr, vstride, stride, form, n, g, c, windows, nullptr,
                         ks, vs, table_stride ? 1u : windows, table_stride, val_offset};
        if (fused) {
            rc = sort_pairs(*D, st, &DA, m, key_bits, &ks, &vs, &launches, nullptr);
        } else {
            rc = launch_decompose(DA, st);
            if (!rc && key_bits) rc = sort_pairs(*D, st, nullptr, m, key_bits, &ks, &vs, &launches, nullptr);
        }
    } else {
        COZK_CUDA(cudaMemcpyAsync(ks, d_keys, m * 4, cudaMemcpyDeviceToDevice, st));
        COZK_CUDA(cudaMemcpyAsync(vs, d_vals, m * 4, cudaMemcpyDeviceToDevice, st));
        rc = sort_pairs(*D, st, nullptr, m, key_bits, &ks, &vs, &launches, nullptr);
    }
    if (rc) return rc;
    COZK_CUDA(cudaMemcpyAsync(d_keys_out, ks, m * 4, cudaMemcpyDeviceToDevice, st));
    COZK_CUDA(cudaMemcpyAsync(d_vals_out, vs, m * 4, cudaMemcpyDeviceToDevice, st));
    COZK_CUDA(cudaStreamSynchronize(st));
    return COZK_OK;
}

// The batched-affine pre-reduction on its own (csrc/affine_kernels.cuh): `rounds` halving rounds over m pairs grouped by key
// (vals index into d_pts).  reference == 0: the engine's batched kernels; != 0: the serial contract kernel (its own inversion per
// addition).  Outputs (device pointers, may be NULL): the reduced list - out_m[0] entries of keys, vals (index into d_pts_out, or the
// skip mark) and points - and per round r the overflow list at d_ovf_keys + r * ovf_stride / d_ovf_pts + r * ovf_stride (64-byte
// points), ovf_counts[r] entries (host array).  out_ms: time of the rounds (CUDA events).
int cozk_test_affine_rounds(cozk_ctx* ctx, int device_index, const void* d_keys, const void* d_vals, size_t m, const void* d_pts,
                            size_t total_buckets, int rounds, int reference, void* d_keys_out, void* d_vals_out, void* d_pts_out,
                            size_t* out_m, void* d_ovf_keys, void* d_ovf_pts, size_t ovf_stride, unsigned* ovf_counts, double* out_ms) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    if (rounds < 1 || rounds > AFF_MAX_ROUNDS || m < 2 || !out_m) return COZK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(D->mu);
    cudaStream_t st = D->stream;
    cudaEvent_t e0, e1;
    COZK_CUDA(cudaEventCreate(&e0));
    COZK_CUDA(cudaEventCreate(&e1));
    AffineStage AS;
    double launches = 0;
    std::vector<void*> tmp;
    auto dalloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes + 256) != cudaSuccess) return nullptr;
        tmp.push_back(p);
        return p;
    };
    COZK_CUDA(cudaEventRecord(e0, st));
    if (!reference) {
        rc = affine_reduce(*D, st, 0, m, total_buckets, (const uint32_t*)d_keys, (const uint32_t*)d_vals, (const affine*)d_pts, rounds,
                           &AS, &launches);
    } else {
        uint32_t* counters = (uint32_t*)dalloc(256);
        if (!counters) return COZK_ERR_CUDA;
        COZK_CUDA(cudaMemsetAsync(counters, 0, 256, st));
        const uint32_t *kin = (const uint32_t*)d_keys, *vin = (const uint32_t*)d_vals;
        const affine* pin = (const affine*)d_pts;
        size_t mr = m;
        AS.rounds = rounds;
        for (int r = 0; r < rounds && !rc; ++r) {
            const size_t mo = affine_round_out(mr), cap = std::min(mo, total_buckets + 1);
            AffineRoundArgs A{};
            A.m = mr;
            A.keys_in = kin;
            A.vals_in = vin;
            A.pts_in = pin;
            A.keys_out = (uint32_t*)dalloc(mo * 4);
            A.vals_out = (uint32_t*)dalloc(mo * 4);
            A.pts_out = (affine*)dalloc(mo * sizeof(affine));
            A.ovf_count = counters + r;
            A.ovf_keys = (uint32_t*)dalloc(cap * 4);
            A.ovf_pts = (affine*)dalloc(cap * sizeof(affine));
            A.ovf_cap = (uint32_t)cap;
            if (!A.keys_out || !A.vals_out || !A.pts_out || !A.ovf_keys || !A.ovf_pts) return COZK_ERR_CUDA;
            rc = affine_round_reference(st, A);
            AS.ovf[r] = OvfAddArgsPub{A.ovf_count, A.ovf_keys, A.ovf_pts, nullptr, A.ovf_cap};
            kin = A.keys_out;
            vin = A.vals_out;
            pin = A.pts_out;
            mr = mo;
        }
        AS.keys = kin;
        AS.vals = vin;
        AS.pts = pin;
        AS.m = mr;
    }
    if (rc) return rc;
    COZK_CUDA(cudaEventRecord(e1, st));
    *out_m = AS.m;
    if (d_keys_out) COZK_CUDA(cudaMemcpyAsync(d_keys_out, AS.keys, AS.m * 4, cudaMemcpyDeviceToDevice, st));
    if (d_vals_out) COZK_CUDA(cudaMemcpyAsync(d_vals_out, AS.vals, AS.m * 4, cudaMemcpyDeviceToDevice, st));
    if (d_pts_out) COZK_CUDA(cudaMemcpyAsync(d_pts_out, AS.pts, AS.m * sizeof(affine), cudaMemcpyDeviceToDevice, st));
    for (int r = 0; r < rounds; ++r) {
        uint32_t cnt = 0;
        COZK_CUDA(cudaMemcpyAsync(&cnt, AS.ovf[r].count, 4, cudaMemcpyDeviceToHost, st));
        COZK_CUDA(cudaStreamSynchronize(st));
        if (cnt > AS.ovf[r].cap || cnt > ovf_stride) {
            set_error("overflow list of a round outgrew its bound");
            return COZK_ERR_INVALID_ARG;
        }
        if (ovf_counts) ovf_counts[r] = cnt;
        if (d_ovf_keys && cnt) COZK_CUDA(cudaMemcpyAsync((uint32_t*)d_ovf_keys + r * ovf_stride, AS.ovf[r].keys, cnt * 4, cudaMemcpyDeviceToDevice, st));
        if (d_ovf_pts && cnt)
            COZK_CUDA(cudaMemcpyAsync((affine*)d_ovf_pts + r * ovf_stride, AS.ovf[r].pts, cnt * sizeof(affine), cudaMemcpyDeviceToDevice, st));
    }
    COZK_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    COZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (out_ms) *out_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (void* p : tmp) cudaFree(p);
    return COZK_OK;
}

int cozk_microbench(cozk_ctx* ctx, int device_index, int which, int blocks, int threads, int iters, double* out_ms,
                    double* out_ops) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    if (!out_ms || !out_ops || blocks < 1 || threads < 32 || threads > 1024 || iters < 1) return COZK_ERR_INVALID_ARG;
    if (which == 3 && threads > 128) threads = 128;
    uint8_t* buf = nullptr;
    COZK_CUDA(cudaMalloc(&buf, 64 * 1024 + 256));
    // inputs: 1024 base points double as field elements (valid curve points keep the madd chain off its exceptional paths)
    k_gen_bases<<<8, 128, 0, D->stream>>>(99, 0, 1024, reinterpret_cast<affine*>(buf));
    COZK_CUDA(cudaGetLastError());
    cudaEvent_t e0, e1;
    COZK_CUDA(cudaEventCreate(&e0));
    COZK_CUDA(cudaEventCreate(&e1));
    double ops = 0;
    for (int rep = 0; rep < 2; ++rep) {  // first repetition is the warm-up
        COZK_CUDA(cudaEventRecord(e0, D->stream));
        if (which == 0) {
            k_bench_imad<<<blocks, threads, 0, D->stream>>>(iters, reinterpret_cast<uint32_t*>(buf + 64 * 1024));
            ops = 8.0 * iters * (double)blocks * threads;
        } else if (which == 1 || which == 2) {
            k_bench_fq<<<blocks, threads, 0, D->stream>>>(which, iters, buf, buf + 64 * 1024);
            ops = 2.0 * iters * (double)blocks * threads;
        } else if (which == 3) {
            k_bench_madd<<<blocks, threads, 0, D->stream>>>(iters, buf, buf + 64 * 1024);
            ops = 2.0 * iters * (double)blocks * threads;
        } else if (which == 4) {
            k_bench_imad_cc<<<blocks, threads, 0, D->stream>>>(iters, reinterpret_cast<uint32_t*>(buf + 64 * 1024));
            ops = 16.0 * iters * (double)blocks * threads;  // 4 chains x 4 wide multiply-adds
        } else if (which == 5 || which == 6) {
            k_bench_imad32<<<blocks, threads, 0, D->stream>>>(which == 6, iters, reinterpret_cast<uint32_t*>(buf + 64 * 1024));
            ops = 8.0 * iters * (double)blocks * threads;
        } else {
            k_bench_fq4<<<blocks, threads, 0, D->stream>>>(iters, buf, buf + 64 * 1024);
            ops = 4.0 * iters * (double)blocks * threads;
        }
        COZK_CUDA(cudaGetLastError());
        COZK_CUDA(cudaEventRecord(e1, D->stream));
        COZK_CUDA(cudaStreamSynchronize(D->stream));
    }
    float ms = 0;
    COZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *out_ms = ms;
    *out_ops = ops;
    return COZK_OK;
}

}  // extern "C"
