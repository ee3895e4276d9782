// Fixed-base batch multiplication: out[i] = scalars[i] * base, affine - the G1 half of SRS generation.
//
// Reference: `BatchMulPreprocessing::new(g, n).batch_mul(&pp_powers)` in ZKMLCommit::setup
// (co-noir-spartan/spartan/src/zk.rs:455-460) and the same step inside ark_poly_commit's MultilinearPC::setup that
// PST13::setup calls (co-jolt/src/poly/commitment/pst13.rs:49-62, 276-279; 18.5 - 25.7 s per run in the reference's
// traces).  SURVEY.md section 8(f), row N3.  The G2 half (powers_of_h) is pairing-side and stays on the host.
//
// Method: 8-bit unsigned windows; table[w][d-1] = d * 2^(8w) * base (32 x 255 affine points, built on the device);
// each scalar is then 32 table look-ups and mixed additions and one inversion - no doublings per scalar.
#include <cstring>

#include "engine.hpp"
#include "msm_kernels.cuh"

namespace cozk {

constexpr uint32_t FB_C = 8, FB_W = 32, FB_ROW = 255;

// thread w: row w of the table in XYZZ, by repeated addition of 2^(8w) * base
__global__ void k_fb_rows(affine base, xyzz* rows) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= FB_W) return;
    xyzz p = xyzz_from_affine(base);
    for (uint32_t k = 0; k < FB_C * w; ++k) p = xyzz_dbl(p);
    xyzz acc = p;
    store_xyzz(&rows[(size_t)w * FB_ROW], acc);
    for (uint32_t d = 2; d <= FB_ROW; ++d) {
        acc = xyzz_add(acc, p);
        store_xyzz(&rows[(size_t)w * FB_ROW + d - 1], acc);
    }
}
// thread e: table entry e to affine (an entry is never the identity: d * 2^(8w) < r)
__global__ void k_fb_affine(const xyzz* rows, affine* table, uint32_t n) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    xyzz p = load_xyzz(&rows[e]);
    fq I = fq_inv(fq_mul(p.ZZ, p.ZZZ));
    store_fq(&table[e].x, fq_mul(p.X, fq_mul(I, p.ZZZ)));
    store_fq(&table[e].y, fq_mul(p.Y, fq_mul(I, p.ZZ)));
}
// thread i: scalar i -> 32 look-ups -> affine wire point
__global__ void __launch_bounds__(128, 4) k_fb_mul(const affine* table, const uint8_t* scalars, size_t stride, int form, size_t n,
                                                   uint8_t* out72) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr s = load_fq(scalars + i * stride);
    if (form == SCALAR_MONT) s = fr_from_mont(s); else s = fr_reduce_canon(s);
    xyzz acc = xyzz_identity();
    for (uint32_t w = 0; w < FB_W; ++w) {
        uint32_t d = (s.v[w >> 2] >> ((w & 3) * 8)) & 0xFFu;
        if (d) acc = xyzz_madd(acc, load_affine(&table[(size_t)w * FB_ROW + d - 1]));
    }
    xyzz_to_wire(acc, out72 + 72 * i);
}

}  // namespace cozk

using namespace cozk;

extern "C" int cozk_fixed_base_batch_mul(cozk_ctx* ctx, const void* base72, const void* scalars, size_t n, size_t stride_bytes,
                                         int form, void* out_points72, cozk_srs* out_srs) {
    if (!ctx || !base72 || (!scalars && n) || (!out_points72 && !out_srs)) {
        set_error("null pointer");
        return COZK_ERR_INVALID_ARG;
    }
    if ((form != COZK_MONT && form != COZK_CANON) || stride_bytes < 32 || (stride_bytes & 15)) {
        set_error("bad scalar form or stride");
        return COZK_ERR_INVALID_ARG;
    }
    const uint8_t* b = reinterpret_cast<const uint8_t*>(base72);
    Device& D = *ctx->devs[0];
    uint8_t *d_sc = nullptr, *d_out = nullptr;
    xyzz* d_rows = nullptr;
    affine* d_table = nullptr;
    auto cleanup = [&]() {
        if (d_sc) cudaFree(d_sc);
        if (d_out) cudaFree(d_out);
        if (d_rows) cudaFree(d_rows);
        if (d_table) cudaFree(d_table);
    };
#define FB_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            set_error(std::string(#call " failed: ") + cudaGetErrorString(e__));   \
            cleanup();                                                             \
            return COZK_ERR_CUDA;                                                  \
        }                                                                          \
    } while (0)
    std::vector<uint8_t> host_out;
    {
        std::lock_guard<std::mutex> lock(D.mu);
        FB_CUDA(cudaSetDevice(D.id));
        FB_CUDA(cudaMalloc(&d_out, std::max<size_t>(n, 1) * 72));
        if (b[64] || n == 0) {
            // multiples of the identity are the identity
            std::vector<uint8_t> ident(std::max<size_t>(n, 1) * 72, 0);
            for (size_t i = 0; i < n; ++i) ident[72 * i + 64] = 1;
            FB_CUDA(cudaMemcpyAsync(d_out, ident.data(), n * 72, cudaMemcpyHostToDevice, D.stream));
            FB_CUDA(cudaStreamSynchronize(D.stream));
        } else {
            affine base;
            memcpy(base.x.v, b, 32);
            memcpy(base.y.v, b + 32, 32);
            size_t sc_bytes = (n - 1) * stride_bytes + 32;
            FB_CUDA(cudaMalloc(&d_sc, sc_bytes));
            FB_CUDA(cudaMalloc(&d_rows, (size_t)FB_W * FB_ROW * sizeof(xyzz)));
            FB_CUDA(cudaMalloc(&d_table, (size_t)FB_W * FB_ROW * sizeof(affine)));
            FB_CUDA(cudaMemcpyAsync(d_sc, scalars, sc_bytes, cudaMemcpyHostToDevice, D.stream));
            k_fb_rows<<<1, 32, 0, D.stream>>>(base, d_rows);
            FB_CUDA(cudaGetLastError());
            k_fb_affine<<<(FB_W * FB_ROW + 63) / 64, 64, 0, D.stream>>>(d_rows, d_table, FB_W * FB_ROW);
            FB_CUDA(cudaGetLastError());
            k_fb_mul<<<(unsigned)((n + 127) / 128), 128, 0, D.stream>>>(d_table, d_sc, stride_bytes, form, n, d_out);
            FB_CUDA(cudaGetLastError());
            FB_CUDA(cudaStreamSynchronize(D.stream));
        }
        // straight into the caller's buffer when there is one (a staging vector would cost a second 72 n byte copy and its page faults)
        if (!out_points72) host_out.resize(n * 72);
        uint8_t* host_dst = out_points72 ? reinterpret_cast<uint8_t*>(out_points72) : host_out.data();
        FB_CUDA(cudaMemcpyAsync(host_dst, d_out, n * 72, cudaMemcpyDeviceToHost, D.stream));
        FB_CUDA(cudaStreamSynchronize(D.stream));
    }
    cleanup();
#undef FB_CUDA
    if (out_srs) {
        // register the points as an SRS (stride 72 with per-point infinity flags, exactly ark_ec's Affine image)
        const uint8_t* pts = out_points72 ? reinterpret_cast<const uint8_t*>(out_points72) : host_out.data();
        std::vector<uint8_t> inf(n);
        for (size_t i = 0; i < n; ++i) inf[i] = pts[72 * i + 64];
        return cozk_srs_register(ctx, pts, n, 72, inf.data(), out_srs);
    }
    return COZK_OK;
}
