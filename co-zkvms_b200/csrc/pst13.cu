// Host-side mirror of the reference's PST13 / MultilinearPC interface for the MSM path (see pst13.hpp), plus the
// two small kernels PST13's opening needs between its MSMs (fold r -> q, r').
#include "pst13.hpp"
#include "../../include/cozk_rep3.h"

#include <cstring>
#include <map>
#include <vector>

#include "engine.hpp"
#include "msm_kernels.cuh"

namespace cozk {

// dense[i] = strided[i * stride]   (share `a` of an AoS Rep3 share array, or a plain copy when stride == 32)
__global__ void k_gather_fr(const uint8_t* src, size_t stride, fr* dst, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    store_fq(&dst[t], load_fq(src + t * stride));
}

// One level of open() (pst13.rs:454-459):  q[b] = r[2b+1] - r[2b];  r'[b] = r[2b]*(1-t) + r[2b+1]*t = r[2b] + t*q[b];
// the MSM scalars are q duplicated: scalars[2b] = scalars[2b+1] = q[b].
// dup == 0 (pair-sum SRS): q is written once per pair.
__global__ void k_open_fold(const fr* r, fr t, fr* q_out, fr* r_next, size_t half, int dup) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= half) return;
    fr lo = load_fq(&r[2 * b]), hi = load_fq(&r[2 * b + 1]);
    fr q = fr_sub(hi, lo);
    if (dup) {
        store_fq(&q_out[2 * b], q);
        store_fq(&q_out[2 * b + 1], q);
    } else {
        store_fq(&q_out[b], q);
    }
    store_fq(&r_next[b], fr_add(lo, fr_mul(t, q)));
}

static int log2_exact(size_t n, unsigned* out) {
    unsigned l = 0;
    while (((size_t)1 << l) < n) ++l;
    if (((size_t)1 << l) != n) return -1;
    *out = l;
    return 0;
}

static void write_commitment(uint8_t* out, uint64_t nv, const uint8_t* point72) {
    memcpy(out, &nv, 8);
    memcpy(out + 8, point72, 72);
}

}  // namespace cozk

using namespace cozk;

extern "C" {

int cozk_pst13_commit(cozk_ctx* ctx, cozk_srs srs, const void* evals, size_t n, size_t stride_bytes, int form,
                      unsigned max_num_bits, void* out_commitment) {
    const void* ptrs[1] = {evals};
    return cozk_pst13_batch_commit(ctx, srs, ptrs, 1, n, stride_bytes, form, max_num_bits ? &max_num_bits : nullptr,
                                   out_commitment);
}

int cozk_pst13_batch_commit(cozk_ctx* ctx, cozk_srs srs, const void* const* polys, size_t k, size_t n, size_t stride_bytes,
                            int form, const unsigned* max_num_bits, void* out_commitments) {
    if (!ctx || !polys || !out_commitments || k == 0) {
        set_error("null pointer or k == 0");
        return COZK_ERR_INVALID_ARG;
    }
    unsigned nv = 0;
    if (n == 0 || log2_exact(n, &nv)) {
        set_error("polynomial length must be a power of two");
        return COZK_ERR_INVALID_ARG;
    }
    // one MSM batch per distinct bit-width hint (the reference's batch_msm dispatches on the scalar variant per polynomial)
    std::map<unsigned, std::vector<size_t>> classes;
    for (size_t j = 0; j < k; ++j) classes[max_num_bits ? max_num_bits[j] : 0].push_back(j);
    uint8_t* out = reinterpret_cast<uint8_t*>(out_commitments);
    for (auto& kv : classes) {
        std::vector<const void*> ptrs;
        for (size_t j : kv.second) ptrs.push_back(polys[j]);
        std::vector<uint8_t> pts(ptrs.size() * 72);
        int rc = msm_dispatch(ctx, -1, srs, 0, n, ptrs.data(), nullptr, ptrs.size(), stride_bytes, form, kv.first, pts.data());
        if (rc) return rc;
        for (size_t i = 0; i < kv.second.size(); ++i) write_commitment(out + COZK_COMMITMENT_BYTES * kv.second[i], nv, &pts[72 * i]);
    }
    return COZK_OK;
}

int cozk_pst13_batch_commit_rep3(cozk_ctx* ctx, cozk_srs srs, const void* const* polys, const uint8_t* is_shared, size_t k,
                                 size_t n, int form, const unsigned* max_num_bits, int commit_to_public,
                                 void* out_commitments, uint8_t* present) {
    if (!ctx || !polys || !is_shared || !out_commitments || !present || k == 0) {
        set_error("null pointer or k == 0");
        return COZK_ERR_INVALID_ARG;
    }
    uint8_t* out = reinterpret_cast<uint8_t*>(out_commitments);
    memset(out, 0, k * COZK_COMMITMENT_BYTES);
    std::vector<const void*> shared, pub;
    std::vector<size_t> shared_idx, pub_idx;
    std::vector<unsigned> pub_bits;
    for (size_t j = 0; j < k; ++j) {
        if (is_shared[j]) {
            shared.push_back(polys[j]);
            shared_idx.push_back(j);
            present[j] = 1;
        } else {
            present[j] = commit_to_public ? 1 : 0;
            if (commit_to_public) {
                pub.push_back(polys[j]);
                pub_idx.push_back(j);
                pub_bits.push_back(max_num_bits ? max_num_bits[j] : 0);
            }
        }
    }
    std::vector<uint8_t> tmp;
    if (!shared.empty()) {
        tmp.resize(shared.size() * COZK_COMMITMENT_BYTES);
        int rc = cozk_pst13_batch_commit(ctx, srs, shared.data(), shared.size(), n, 64, form, nullptr, tmp.data());
        if (rc) return rc;
        for (size_t i = 0; i < shared.size(); ++i)
            memcpy(out + COZK_COMMITMENT_BYTES * shared_idx[i], &tmp[COZK_COMMITMENT_BYTES * i], COZK_COMMITMENT_BYTES);
    }
    if (!pub.empty()) {
        tmp.resize(pub.size() * COZK_COMMITMENT_BYTES);
        int rc = cozk_pst13_batch_commit(ctx, srs, pub.data(), pub.size(), n, 32, form, pub_bits.data(), tmp.data());
        if (rc) return rc;
        for (size_t i = 0; i < pub.size(); ++i)
            memcpy(out + COZK_COMMITMENT_BYTES * pub_idx[i], &tmp[COZK_COMMITMENT_BYTES * i], COZK_COMMITMENT_BYTES);
    }
    return COZK_OK;
}

static int open_check_levels(cozk_ctx* ctx, const cozk_srs* level_srs, const cozk_srs* level_pairs, size_t nv) {
    // assert_eq!(nv, ck.nv): every level must hold exactly 2^(nv-i) points (pair sums: half of that)
    for (size_t i = 0; i < nv; ++i) {
        size_t len = 0;
        int rc = cozk_srs_len(ctx, level_srs[i], &len);
        if (rc) return rc;
        if (len != ((size_t)1 << (nv - i))) {
            set_error("Invalid size of polynomial: SRS level length does not match nv");
            return COZK_ERR_KEY_LENGTH;
        }
        if (level_pairs) {
            rc = cozk_srs_len(ctx, level_pairs[i], &len);
            if (rc) return rc;
            if (len != ((size_t)1 << (nv - i - 1))) {
                set_error("pair-sum SRS level length does not match nv");
                return COZK_ERR_KEY_LENGTH;
            }
        }
    }
    return COZK_OK;
}

int cozk_pst13_open(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv, const void* evals, size_t stride_bytes,
                    const void* point, int form, void* out_proofs, void* out_eval) {
    return cozk_pst13_open_paired(ctx, level_srs, nullptr, nv, evals, stride_bytes, point, form, out_proofs, out_eval);
}

int cozk_pst13_open_paired(cozk_ctx* ctx, const cozk_srs* level_srs, const cozk_srs* level_pairs, size_t nv,
                           const void* evals, size_t stride_bytes, const void* point, int form, void* out_proofs,
                           void* out_eval) {
    if (!ctx || !level_srs || !evals || !point || !out_proofs || !out_eval || nv == 0 || nv > 30) {
        set_error("null pointer or bad nv");
        return COZK_ERR_INVALID_ARG;
    }
    if (form != COZK_MONT || stride_bytes < 32 || (stride_bytes & 15)) {
        set_error("open() takes Montgomery-form Fr values at a stride that is a multiple of 16");
        return COZK_ERR_INVALID_ARG;
    }
    int rc = open_check_levels(ctx, level_srs, level_pairs, nv);
    if (rc) return rc;
    Device& D = *ctx->devs[0];
    size_t n = (size_t)1 << nv;
    uint8_t* d_in = nullptr;
    fr* d_r0 = nullptr;
    {
        std::lock_guard<std::mutex> lock(D.mu);
        cudaError_t e = cudaSetDevice(D.id);
        size_t in_bytes = (n - 1) * stride_bytes + 32;
        if (e == cudaSuccess) e = cudaMalloc(&d_in, in_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&d_r0, n * sizeof(fr));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, evals, in_bytes, cudaMemcpyHostToDevice, D.stream);
        if (e == cudaSuccess) {
            k_gather_fr<<<(unsigned)((n + 255) / 256), 256, 0, D.stream>>>(d_in, stride_bytes, d_r0, n);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
        if (d_in) cudaFree(d_in);
        if (e != cudaSuccess) {
            if (d_r0) cudaFree(d_r0);
            set_error(std::string("open(): staging the evaluations failed: ") + cudaGetErrorString(e));
            return COZK_ERR_CUDA;
        }
    }
    return pst13_open_device(ctx, level_srs, level_pairs, nv, d_r0, point, out_proofs, out_eval);
}

}  // extern "C"

namespace cozk {

// d_r0: 2^nv dense Montgomery evaluations on device 0, owned by this call from here on.
int pst13_open_device(cozk_ctx* ctx, const cozk_srs* level_srs, const cozk_srs* level_pairs, size_t nv, fr* d_r0,
                      const void* point, void* out_proofs, void* out_eval) {
    Device& D = *ctx->devs[0];
    size_t n = (size_t)1 << nv;
    fr *d_r[2] = {d_r0, nullptr}, *d_q = nullptr;
    int rc = COZK_OK;
    auto cleanup = [&]() {
        cudaSetDevice(D.id);
        if (d_r[0]) cudaFree(d_r[0]);
        if (d_r[1]) cudaFree(d_r[1]);
        if (d_q) cudaFree(d_q);
    };
#define OPEN_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            set_error(std::string(#call " failed: ") + cudaGetErrorString(e__));         \
            cleanup();                                                                   \
            return COZK_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)
    {
        std::lock_guard<std::mutex> lock(D.mu);
        OPEN_CUDA(cudaSetDevice(D.id));
        OPEN_CUDA(cudaMalloc(&d_r[1], (n / 2 + 1) * sizeof(fr)));
        OPEN_CUDA(cudaMalloc(&d_q, (level_pairs ? n / 2 + 1 : n) * sizeof(fr)));
    }
    const uint8_t* pt = reinterpret_cast<const uint8_t*>(point);
    uint8_t* proofs = reinterpret_cast<uint8_t*>(out_proofs);
    int cur = 0;
    for (size_t i = 0; i < nv; ++i) {
        size_t k = nv - i, half = (size_t)1 << (k - 1);
        fr t;
        memcpy(t.v, pt + 32 * i, 32);
        {
            std::lock_guard<std::mutex> lock(D.mu);
            OPEN_CUDA(cudaSetDevice(D.id));
            k_open_fold<<<(unsigned)((half + 127) / 128), 128, 0, D.stream>>>(d_r[cur], t, d_q, d_r[cur ^ 1], half,
                                                                             level_pairs ? 0 : 1);
            OPEN_CUDA(cudaGetLastError());
            OPEN_CUDA(cudaStreamSynchronize(D.stream));
        }
        const void* vec[1] = {d_q};
        // with pair sums: sum_b q[b] * (P[2b] + P[2b+1]) over `half` points; without: q duplicated over 2 * half points
        if (level_pairs)
            rc = msm_dispatch(ctx, 0, level_pairs[i], 0, half, nullptr, vec, 1, 32, COZK_MONT, 0, proofs + 72 * i);
        else
            rc = msm_dispatch(ctx, 0, level_srs[i], 0, 2 * half, nullptr, vec, 1, 32, COZK_MONT, 0, proofs + 72 * i);
        if (rc) {
            cleanup();
            return rc;
        }
        cur ^= 1;
    }
    {
        std::lock_guard<std::mutex> lock(D.mu);
        OPEN_CUDA(cudaSetDevice(D.id));
        OPEN_CUDA(cudaMemcpy(out_eval, d_r[cur], 32, cudaMemcpyDeviceToHost));
    }
    cleanup();
#undef OPEN_CUDA
    return COZK_OK;
}

}  // namespace cozk

extern "C" {

int cozk_pst13_combine_commitment_shares(const void* commitments, size_t count, void* out_commitment) {
    if (!commitments || !out_commitment || count == 0) {
        set_error("null pointer or no commitments");
        return COZK_ERR_INVALID_ARG;
    }
    const uint8_t* in = reinterpret_cast<const uint8_t*>(commitments);
    uint64_t nv = 0;
    memcpy(&nv, in, 8);
    std::vector<uint8_t> pts(count * 72);
    for (size_t i = 0; i < count; ++i) {
        uint64_t nvi;
        memcpy(&nvi, in + COZK_COMMITMENT_BYTES * i, 8);
        if (nvi != nv) {
            set_error("commitment shares disagree on nv");
            return COZK_ERR_INVALID_ARG;
        }
        memcpy(&pts[72 * i], in + COZK_COMMITMENT_BYTES * i + 8, 72);
    }
    uint8_t sum[72];
    int rc = cozk_g1_sum(pts.data(), count, sum);
    if (rc) return rc;
    write_commitment(reinterpret_cast<uint8_t*>(out_commitment), nv, sum);
    return COZK_OK;
}

int cozk_pst13_coordinate_prove(const void* proofs, size_t parties, size_t len, void* out_proofs) {
    if (!proofs || !out_proofs || parties == 0) {
        set_error("null pointer or no parties");
        return COZK_ERR_INVALID_ARG;
    }
    const uint8_t* in = reinterpret_cast<const uint8_t*>(proofs);
    uint8_t* out = reinterpret_cast<uint8_t*>(out_proofs);
    std::vector<uint8_t> col(parties * 72);
    for (size_t i = 0; i < len; ++i) {
        for (size_t p = 0; p < parties; ++p) memcpy(&col[72 * p], in + (p * len + i) * 72, 72);
        int rc = cozk_g1_sum(col.data(), parties, out + 72 * i);
        if (rc) return rc;
    }
    return COZK_OK;
}

int cozk_combine_comm(const void* commitments, size_t count, void* out_commitment) {
    unsigned lg = 0;
    if (!commitments || !out_commitment || count == 0 || log2_exact(count, &lg)) {
        set_error("null pointer, or the number of chunk commitments is not a power of two");
        return COZK_ERR_INVALID_ARG;
    }
    const uint8_t* in = reinterpret_cast<const uint8_t*>(commitments);
    uint64_t nv = 0;
    memcpy(&nv, in, 8);
    std::vector<uint8_t> pts(count * 72);
    for (size_t i = 0; i < count; ++i) memcpy(&pts[72 * i], in + COZK_COMMITMENT_BYTES * i + 8, 72);
    uint8_t sum[72];
    int rc = cozk_g1_sum(pts.data(), count, sum);
    if (rc) return rc;
    write_commitment(reinterpret_cast<uint8_t*>(out_commitment), nv + lg, sum);
    return COZK_OK;
}

}  // extern "C"
