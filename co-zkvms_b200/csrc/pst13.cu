// Host-side mirror of the reference's PST13 / MultilinearPC interface for the MSM path (see include/cozk_pst13.h), plus the
// two small kernels PST13's opening needs between its MSMs (fold r -> q, r').
#include "../../include/cozk_pst13.h"
#include "../../include/cozk_rep3.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "engine.hpp"
#include "msm_plan.hpp"
#include "rep3_kernels.cuh"

namespace cozk {

// dense[i] = strided[i * stride]   (share `a` of an AoS Rep3 share array, or a plain copy when stride == 32)
// canon: the source holds canonical integers (a widened small-scalar polynomial): converted to Montgomery form on the way
__global__ void k_gather_fr(const uint8_t* src, size_t stride, fr* dst, size_t n, int canon) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fr v = load_fq(src + t * stride);
    if (canon) v = fr_mont_from_canon(v);
    store_fq(&dst[t], v);
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

static int open_check_levels(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv) {
    // assert_eq!(nv, ck.nv): every level must hold exactly 2^(nv-i) points
    for (size_t i = 0; i < nv; ++i) {
        size_t len = 0;
        int rc = cozk_srs_len(ctx, level_srs[i], &len);
        if (rc) return rc;
        if (len != ((size_t)1 << (nv - i))) {
            set_error("Invalid size of polynomial: SRS level length does not match nv");
            return COZK_ERR_KEY_LENGTH;
        }
    }
    return COZK_OK;
}

static int open_check_args(cozk_ctx* ctx, const void* evals, const void* point, void* out_proofs, void* out_eval, size_t nv,
                           size_t stride_bytes, int form) {
    if (!ctx || !evals || !point || !out_proofs || !out_eval || nv == 0 || nv > 30) {
        set_error("null pointer or bad nv");
        return COZK_ERR_INVALID_ARG;
    }
    if (form != COZK_MONT || stride_bytes < 32 || (stride_bytes & 15)) {
        set_error("open() takes Montgomery-form Fr values at a stride that is a multiple of 16");
        return COZK_ERR_INVALID_ARG;
    }
    return COZK_OK;
}

int cozk_pst13_batch_commit_packed(cozk_ctx* ctx, cozk_srs srs, const void* const* polys, const int* kinds, size_t k, size_t n,
                                   int commit_to_public, void* out_commitments, uint8_t* present) {
    if (!ctx || !polys || !kinds || !out_commitments || !present || k == 0) {
        set_error("null pointer or k == 0");
        return COZK_ERR_INVALID_ARG;
    }
    // The packed image goes to the device as it is (cozk_poly_upload widens the small kinds there), the commitments run
    // over the resident polynomials, nothing stays behind.  Public polynomials nobody keeps are not even uploaded.
    const int nd = cozk_device_count(ctx);
    std::vector<cozk_poly> handles;
    std::vector<size_t> index;
    uint8_t* o = reinterpret_cast<uint8_t*>(out_commitments);
    int rc = COZK_OK;
    for (size_t j = 0; j < k && !rc; ++j) {
        present[j] = 0;
        memset(o + COZK_COMMITMENT_BYTES * j, 0, COZK_COMMITMENT_BYTES);
        if (!polys[j]) {
            set_error("null polynomial");
            rc = COZK_ERR_INVALID_ARG;
            break;
        }
        if (kinds[j] != COZK_POLY_SHARED && !commit_to_public) continue;  // MaybeShared::Public(None)
        cozk_poly h = 0;
        rc = cozk_poly_upload(ctx, (int)(handles.size() % (size_t)nd), polys[j], n, kinds[j], &h);
        if (!rc) {
            handles.push_back(h);
            index.push_back(j);
        }
    }
    if (!rc && !handles.empty()) {
        std::vector<uint8_t> comm(handles.size() * COZK_COMMITMENT_BYTES), pres(handles.size());
        rc = cozk_pst13_batch_commit_polys(ctx, srs, handles.data(), handles.size(), 1, comm.data(), pres.data());
        for (size_t q = 0; q < handles.size() && !rc; ++q) {
            present[index[q]] = pres[q];
            memcpy(o + COZK_COMMITMENT_BYTES * index[q], &comm[q * COZK_COMMITMENT_BYTES], COZK_COMMITMENT_BYTES);
        }
    }
    const std::string keep = rc ? cozk_last_error() : "";
    for (cozk_poly h : handles) cozk_poly_release(ctx, h);
    if (rc) set_error(keep);
    return rc;
}

int cozk_pst13_open(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv, const void* evals, size_t stride_bytes,
                    const void* point, int form, void* out_proofs, void* out_eval) {
    int rc = open_check_args(ctx, evals, point, out_proofs, out_eval, nv, stride_bytes, form);
    if (rc) return rc;
    if (!level_srs) {
        set_error("null pointer");
        return COZK_ERR_INVALID_ARG;
    }
    rc = open_check_levels(ctx, level_srs, nv);
    if (rc) return rc;
    OpenSource src;
    src.host = evals;
    src.stride = stride_bytes;
    return pst13_open_device(ctx, level_srs, nullptr, nv, src, point, out_proofs, out_eval);
}

int cozk_pst13_open_keyed(cozk_ctx* ctx, cozk_open_key key, const void* evals, size_t stride_bytes, const void* point,
                          int form, void* out_proofs, void* out_eval) {
    OpenKey K;
    if (!ctx) return COZK_ERR_INVALID_ARG;
    int rc = open_key_lookup(ctx, key, &K);
    if (rc) return rc;
    rc = open_check_args(ctx, evals, point, out_proofs, out_eval, K.nv, stride_bytes, form);
    if (rc) return rc;
    OpenSource src;
    src.host = evals;
    src.stride = stride_bytes;
    return pst13_open_device(ctx, K.level_srs.data(), &K, K.nv, src, point, out_proofs, out_eval);
}

int cozk_pst13_open_key_create(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv, cozk_open_key* out) {
    if (!ctx || !level_srs || !out || nv == 0 || nv > 30) {
        set_error("null pointer or bad nv");
        return COZK_ERR_INVALID_ARG;
    }
    int rc = open_check_levels(ctx, level_srs, nv);
    if (rc) return rc;
    OpenKey K;
    K.nv = nv;
    K.level_srs.assign(level_srs, level_srs + nv);
    size_t small_log2;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        small_log2 = (size_t)ctx->opt_open_small_log2;
    }
    // up to nv = 20 every level goes into the one batch of "small" levels (measured, keyed opening at nv = 18 / 20: 2.04 / 4.42 ms
    // with two batches, 1.76 / 4.30 ms with one; at nv = 22 two batches win, 12.9 against 13.2 ms)
    // (option "open_one_batch_max_nv", default 20; 0 = always split by "open_small_log2")
    if (small_log2 != 0 && nv <= (size_t)ctx->opt_open_one_batch_max_nv && small_log2 + 1 < nv) small_log2 = nv - 1;
    // level i has 2^(nv-1-i) quotient scalars; the levels with at most 2^small_log2 of them are batched (none if 0)
    K.first_small = small_log2 == 0 ? nv : (nv > small_log2 + 1 ? nv - small_log2 - 1 : 0);
    K.small_n = K.first_small < nv ? ((size_t)1 << (nv - K.first_small)) - 1 : 0;
    Device& D = *ctx->devs[0];
    // two concatenated SRSs of pair sums: the big levels (opened by ONE ragged batch: vectors of 2^(nv-1), 2^(nv-2), ..
    // scalars through one decompose / sort / accumulate / reduce) and the small ones (one batched MSM, zero-padded vectors)
    size_t big_n = 0;
    for (size_t i = 0; i < K.first_small; ++i) {
        K.big_off.push_back(big_n);
        big_n += (size_t)1 << (nv - 1 - i);
    }
    affine *d_small = nullptr, *d_big = nullptr;
    uint8_t *d_small_inf = nullptr, *d_big_inf = nullptr;
    auto fail = [&](int code) {
        if (K.big_srs) cozk_srs_release(ctx, K.big_srs);
        cudaSetDevice(D.id);
        for (void* p : {(void*)d_small, (void*)d_small_inf, (void*)d_big, (void*)d_big_inf})
            if (p) cudaFree(p);
        return code;
    };
    {
        std::lock_guard<std::mutex> lock(D.mu);
        cudaError_t e = cudaSetDevice(D.id);
        if (e == cudaSuccess && K.small_n) e = cudaMalloc(&d_small, K.small_n * sizeof(affine));
        if (e == cudaSuccess && K.small_n) e = cudaMalloc(&d_small_inf, K.small_n);
        if (e == cudaSuccess && big_n) e = cudaMalloc(&d_big, big_n * sizeof(affine));
        if (e == cudaSuccess && big_n) e = cudaMalloc(&d_big_inf, big_n);
        if (e != cudaSuccess) {
            set_error(std::string("open key: allocation failed: ") + cudaGetErrorString(e));
            return fail(COZK_ERR_CUDA);
        }
    }
    size_t off = 0;
    for (size_t i = 0; i < nv; ++i) {
        size_t half = (size_t)1 << (nv - 1 - i);
        size_t got = 0;
        if (i < K.first_small) {
            rc = srs_pair_sums_into(ctx, level_srs[i], d_big + K.big_off[i], d_big_inf + K.big_off[i], &got);
            if (rc) return fail(rc);
        } else {
            rc = srs_pair_sums_into(ctx, level_srs[i], d_small + off, d_small_inf + off, &got);
            if (rc) return fail(rc);
            K.small_off.push_back(off);
            off += half;
        }
    }
    if (big_n) {
        // window of the big SRS's table: all levels share it, every level has its own bucket set:
        // cost = 11 * W(c) * (points of all levels) + 45 * levels * 2^(c-1)
        uint32_t best_c = 0;
        double best = 0;
        for (uint32_t c = 10; c <= C_MAX; ++c) {
            const double W = (double)windows_for(254, c);
            if (W * (double)big_n >= 2147483647.0) continue;
            const double cost = 11.0 * W * (double)big_n + 45.0 * (double)K.first_small * (double)(1u << (c - 1));
            if (!best_c || cost < best) {
                best = cost;
                best_c = c;
            }
        }
        rc = srs_register_from_device(ctx, 0, d_big, d_big_inf, big_n, &K.big_srs, 0, best_c);  // read by device 0 only
        if (rc) return fail(rc);
    }
    if (K.small_n) {
        // the same rule for the small levels (one ragged batch, a bucket set per level); option "open_small_window" overrides
        const size_t g_small = nv - K.first_small;
        uint32_t small_c = (uint32_t)ctx->opt_open_small_window;
        if (!small_c) {
            double best = 0;
            for (uint32_t c = 8; c <= C_MAX; ++c) {
                const double cost = 11.0 * (double)windows_for(254, c) * (double)K.small_n + 45.0 * (double)g_small * (double)(1u << (c - 1));
                if (!small_c || cost < best) {
                    best = cost;
                    small_c = c;
                }
            }
        }
        rc = srs_register_from_device(ctx, 0, d_small, d_small_inf, K.small_n, &K.small_srs, 0, small_c);  // read by device 0 only
        if (rc) return fail(rc);
    }
    {
        std::lock_guard<std::mutex> lock(D.mu);
        cudaSetDevice(D.id);
        for (void* p : {(void*)d_small, (void*)d_small_inf, (void*)d_big, (void*)d_big_inf})
            if (p) cudaFree(p);
    }
    std::lock_guard<std::mutex> lock(ctx->mu);
    *out = ctx->next_handle++;
    ctx->open_keys[*out] = K;
    return COZK_OK;
}

int cozk_pst13_open_key_release(cozk_ctx* ctx, cozk_open_key key) {
    if (!ctx) return COZK_ERR_INVALID_ARG;
    OpenKey K;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        auto it = ctx->open_keys.find(key);
        if (it == ctx->open_keys.end()) {
            set_error("unknown opening key");
            return COZK_ERR_BAD_HANDLE;
        }
        K = it->second;
        ctx->open_keys.erase(it);
    }
    if (K.big_srs) cozk_srs_release(ctx, K.big_srs);
    if (K.small_srs) cozk_srs_release(ctx, K.small_srs);
    return COZK_OK;
}

}  // extern "C"

namespace cozk {

int open_key_lookup(cozk_ctx* ctx, uint64_t h, OpenKey* out) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->open_keys.find(h);
    if (it == ctx->open_keys.end()) {
        set_error("unknown opening key");
        return COZK_ERR_BAD_HANDLE;
    }
    *out = it->second;
    return COZK_OK;
}

// The evaluations are brought into a dense Montgomery vector on device 0 (copy_share_a, dense_mlpoly.rs:102-110, as a strided
// read).  All nv folds run back to back on the stream (pst13.rs:454-458); the MSMs follow: without a key one per level over
// the duplicated quotient scalars (the reference's schedule, pst13.rs:459-469); with a key one per large level over its pair
// sums and ONE batched MSM for all small levels - level j's scalar vector is zero outside its own slice of the concatenated
// small SRS, so the batch returns every level's sum separately.  Scratch memory is the device's grow-only open_* buffers.
int pst13_open_device(cozk_ctx* ctx, const cozk_srs* level_srs, const OpenKey* key, size_t nv, const OpenSource& src,
                      const void* point, void* out_proofs, void* out_eval) {
    Device& D = *ctx->devs[0];
    std::lock_guard<std::mutex> open_lock(D.open_mu);
    size_t n = (size_t)1 << nv;
    auto cleanup = [&]() {};
#define OPEN_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            set_error(std::string(#call " failed: ") + cudaGetErrorString(e__));         \
            cleanup();                                                                   \
            return COZK_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)
    const size_t first_small = key ? key->first_small : nv;
    const size_t g_small = nv - first_small;
    // quotient vectors of the individually opened levels, one after the other: level i at q_off[i]
    std::vector<size_t> q_off(nv, 0);
    size_t q_total = 0;
    for (size_t i = 0; i < first_small; ++i) {
        q_off[i] = q_total;
        q_total += key ? ((size_t)1 << (nv - 1 - i)) : ((size_t)1 << (nv - i));
    }
    const uint8_t* pt = reinterpret_cast<const uint8_t*>(point);
    uint8_t* proofs = reinterpret_cast<uint8_t*>(out_proofs);
    int cur = 0;
    fr *d_r[2] = {nullptr, nullptr}, *d_q = nullptr, *d_qs = nullptr;
    {
        std::lock_guard<std::mutex> lock(D.mu);
        OPEN_CUDA(cudaSetDevice(D.id));
        int rc0 = D.open_r[0].ensure(n * sizeof(fr));
        if (!rc0) rc0 = D.open_r[1].ensure((n / 2 + 1) * sizeof(fr));
        if (!rc0) rc0 = D.open_q.ensure(std::max<size_t>(q_total, 1) * sizeof(fr));
        if (!rc0 && g_small) rc0 = D.open_qs.ensure(g_small * key->small_n * sizeof(fr));
        if (!rc0 && src.host) rc0 = D.open_in.ensure((n - 1) * src.stride + 32);
        if (rc0) return rc0;
        d_r[0] = D.open_r[0].as<fr>();
        d_r[1] = D.open_r[1].as<fr>();
        d_q = D.open_q.as<fr>();
        d_qs = D.open_qs.as<fr>();
        const uint8_t* from = src.dev;
        if (src.host) {
            OPEN_CUDA(cudaMemcpyAsync(D.open_in.p, src.host, (n - 1) * src.stride + 32, cudaMemcpyHostToDevice, D.stream));
            from = D.open_in.as<uint8_t>();
        }
        k_gather_fr<<<(unsigned)((n + 255) / 256), 256, 0, D.stream>>>(from, src.stride, d_r[0], n, src.canon);
        OPEN_CUDA(cudaGetLastError());
        if (g_small) OPEN_CUDA(cudaMemsetAsync(d_qs, 0, g_small * key->small_n * sizeof(fr), D.stream));
        for (size_t i = 0; i < nv; ++i) {
            size_t half = (size_t)1 << (nv - 1 - i);
            fr t;
            memcpy(t.v, pt + 32 * i, 32);
            fr* q_out = i < first_small ? d_q + q_off[i] : d_qs + (i - first_small) * key->small_n + key->small_off[i - first_small];
            k_open_fold<<<(unsigned)((half + 127) / 128), 128, 0, D.stream>>>(d_r[cur], t, q_out, d_r[cur ^ 1], half, key ? 0 : 1);
            OPEN_CUDA(cudaGetLastError());
            cur ^= 1;
        }
        OPEN_CUDA(cudaMemcpyAsync(out_eval, d_r[cur], 32, cudaMemcpyDeviceToHost, D.stream));
        OPEN_CUDA(cudaStreamSynchronize(D.stream));
    }
    int rc = COZK_OK;
    const bool trace = getenv("COZK_OPEN_TRACE") != nullptr;  // per-stage wall times on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    auto t_start = now();
    bool big_done = false;
    if (key && first_small) {
        // all big levels in one ragged batch against the concatenated pair-sum SRS
        auto t0 = now();
        std::vector<size_t> offs(first_small), lens(first_small);
        std::vector<const void*> vecs(first_small);
        for (size_t i = 0; i < first_small; ++i) {
            offs[i] = key->big_off[i];
            lens[i] = (size_t)1 << (nv - 1 - i);
            vecs[i] = d_q + q_off[i];
        }
        const int rrc = msm_ragged_device(ctx, 0, key->big_srs, offs.data(), lens.data(), vecs.data(), first_small, 32, COZK_MONT, proofs);
        big_done = rrc == COZK_OK;
        if (rrc != COZK_OK && rrc != COZK_ERR_INVALID_ARG) rc = rrc;  // over budget: fall back to one call per level below
        if (trace) fprintf(stderr, "[open] %zu big levels in one ragged batch: %.3f ms (rc %d)\n", first_small, ms_since(t0), rrc);
    }
    for (size_t i = 0; i < first_small && !rc && !big_done; ++i) {
        auto t0 = now();
        const void* vec[1] = {d_q + q_off[i]};
        if (key)
            rc = msm_dispatch(ctx, 0, key->big_srs, key->big_off[i], (size_t)1 << (nv - 1 - i), nullptr, vec, 1, 32, COZK_MONT, 0, proofs + 72 * i);
        else
            rc = msm_dispatch(ctx, 0, level_srs[i], 0, (size_t)1 << (nv - i), nullptr, vec, 1, 32, COZK_MONT, 0, proofs + 72 * i);
        if (trace) fprintf(stderr, "[open] level %zu: %.3f ms\n", i, ms_since(t0));
    }
    bool small_done = false;
    if (!rc && g_small && ctx->opt_open_small_ragged) {
        // the small levels as ONE ragged batch too: only their 2^(nv - first_small) - 1 real scalars are decomposed and sorted
        // (the zero-padded batch below makes g_small times as many pairs, nearly all of them skipped)
        auto t0 = now();
        std::vector<size_t> offs(g_small), lens(g_small);
        std::vector<const void*> vecs(g_small);
        for (size_t j = 0; j < g_small; ++j) {
            offs[j] = key->small_off[j];
            lens[j] = (size_t)1 << (nv - 1 - (first_small + j));
            vecs[j] = d_qs + j * key->small_n + key->small_off[j];
        }
        const int rrc = msm_ragged_device(ctx, 0, key->small_srs, offs.data(), lens.data(), vecs.data(), g_small, 32, COZK_MONT,
                                          proofs + 72 * first_small);
        small_done = rrc == COZK_OK;
        if (rrc != COZK_OK && rrc != COZK_ERR_INVALID_ARG) rc = rrc;
        if (trace) fprintf(stderr, "[open] %zu small levels in one ragged batch: %.3f ms (rc %d)\n", g_small, ms_since(t0), rrc);
    }
    if (!rc && g_small && !small_done) {
        auto t0 = now();
        std::vector<const void*> vecs(g_small);
        for (size_t j = 0; j < g_small; ++j) vecs[j] = d_qs + j * key->small_n;
        rc = msm_dispatch(ctx, 0, key->small_srs, 0, key->small_n, nullptr, vecs.data(), g_small, 32, COZK_MONT, 0,
                          proofs + 72 * first_small);
        if (trace) fprintf(stderr, "[open] %zu small levels in one batch of %zu points: %.3f ms\n", g_small, key->small_n, ms_since(t0));
    }
    if (trace) fprintf(stderr, "[open] MSMs %.3f ms\n", ms_since(t_start));
#undef OPEN_CUDA
    return rc;
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
