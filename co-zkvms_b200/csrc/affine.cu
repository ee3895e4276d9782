// Host driver of the batched-affine pre-reduction (affine_kernels.cuh).
#include "affine_kernels.cuh"
#include "engine.hpp"

#include <algorithm>

namespace cozk {

__global__ void __launch_bounds__(64) k_ovf_add(OvfAddArgs A) { ovf_add_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }

int affine_setup_device() {
    COZK_CUDA(cudaFuncSetAttribute(k_affine_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AffineSmem)));
    return COZK_OK;
}

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

// R rounds over the m sorted pairs (keys, vals -> bases) on stream st; scratch set `parity` of the device.  On return
// *out describes the reduced list (what the accumulate levels now run over) and the overflow lists of the rounds.
int affine_reduce(Device& D, cudaStream_t st, int parity, size_t m, size_t total_buckets, const uint32_t* keys, const uint32_t* vals,
                  const affine* bases, int rounds, AffineStage* out, double* launches) {
    if (rounds < 1 || rounds > AFF_MAX_ROUNDS || m < 2) {
        set_error("internal: affine_reduce called with nothing to do");
        return COZK_ERR_INVALID_ARG;
    }
    // carve one scratch buffer: counters | per round: keys, vals, points, overflow keys, overflow points | batch products, inverses
    size_t mr[AFF_MAX_ROUNDS + 1], cap[AFF_MAX_ROUNDS];
    mr[0] = m;
    size_t bytes = 256;
    for (int r = 0; r < rounds; ++r) {
        mr[r + 1] = affine_round_out(mr[r]);
        cap[r] = std::min(mr[r + 1], total_buckets + 1);
        bytes += 2 * up256(mr[r + 1] * 4) + up256(mr[r + 1] * sizeof(affine)) + up256(cap[r] * 4) + up256(cap[r] * sizeof(affine));
    }
    const size_t max_batches = (mr[1] + AFF_BATCH - 1) / AFF_BATCH;
    bytes += 2 * up256(max_batches * sizeof(fq));
    int rc = D.aff[parity].ensure(bytes);
    if (rc) return rc;
    uint8_t* p = D.aff[parity].as<uint8_t>();
    uint32_t* counters = reinterpret_cast<uint32_t*>(p);
    p += 256;
    COZK_CUDA(cudaMemsetAsync(counters, 0, 256, st));
    fq* prod = reinterpret_cast<fq*>(p);
    p += up256(max_batches * sizeof(fq));
    fq* inv = reinterpret_cast<fq*>(p);
    p += up256(max_batches * sizeof(fq));
    const uint32_t *kin = keys, *vin = vals;
    const affine* pin = bases;
    out->rounds = rounds;
    for (int r = 0; r < rounds; ++r) {
        AffineRoundArgs A{};
        A.m = mr[r];
        A.keys_in = kin;
        A.vals_in = vin;
        A.pts_in = pin;
        A.keys_out = reinterpret_cast<uint32_t*>(p);
        p += up256(mr[r + 1] * 4);
        A.vals_out = reinterpret_cast<uint32_t*>(p);
        p += up256(mr[r + 1] * 4);
        A.pts_out = reinterpret_cast<affine*>(p);
        p += up256(mr[r + 1] * sizeof(affine));
        A.ovf_count = counters + r;
        A.ovf_keys = reinterpret_cast<uint32_t*>(p);
        p += up256(cap[r] * 4);
        A.ovf_pts = reinterpret_cast<affine*>(p);
        p += up256(cap[r] * sizeof(affine));
        A.ovf_cap = (uint32_t)cap[r];
        A.batch_prod = prod;
        A.batch_inv = inv;
        const unsigned batches = (unsigned)((mr[r + 1] + AFF_BATCH - 1) / AFF_BATCH);
        k_affine_prod<<<batches, AFF_T, 0, st>>>(A);
        COZK_CUDA(cudaGetLastError());
        k_affine_inv<<<(batches + 127) / 128, 128, 0, st>>>(prod, inv, batches);
        COZK_CUDA(cudaGetLastError());
        k_affine_apply<<<batches, AFF_T, sizeof(AffineSmem), st>>>(A);
        COZK_CUDA(cudaGetLastError());
        *launches += 3;
        out->ovf[r] = OvfAddArgsPub{A.ovf_count, A.ovf_keys, A.ovf_pts, nullptr, A.ovf_cap};
        kin = A.keys_out;
        vin = A.vals_out;
        pin = A.pts_out;
    }
    out->keys = kin;
    out->vals = vin;
    out->pts = pin;
    out->m = mr[rounds];
    return COZK_OK;
}

// the overflow lists join the bucket set, round after round (distinct keys inside a list; the lists one after the other)
int affine_overflow_adds(cudaStream_t st, const AffineStage& S, xyzz* buckets, double* launches) {
    for (int r = 0; r < S.rounds; ++r) {
        const OvfAddArgs A{S.ovf[r].count, S.ovf[r].keys, S.ovf[r].pts, buckets, S.ovf[r].cap};
        if (A.cap == 0) continue;
        k_ovf_add<<<(A.cap + 63) / 64, 64, 0, st>>>(A);
        COZK_CUDA(cudaGetLastError());
        *launches += 1;
    }
    return COZK_OK;
}

// one round through the serial contract kernel (test entry points only)
int affine_round_reference(cudaStream_t st, const AffineRoundArgs& A) {
    const size_t mout = affine_round_out(A.m);
    k_affine_ref<<<(unsigned)((mout + 127) / 128), 128, 0, st>>>(A);
    COZK_CUDA(cudaGetLastError());
    return COZK_OK;
}

}  // namespace cozk
