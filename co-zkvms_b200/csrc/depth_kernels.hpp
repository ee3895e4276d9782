// Launchers of the depth-bound stages of the MSM pipeline (depth_kernels.cu), called from msm.cu's run_group.
#pragma once
#include <cuda_runtime.h>

#include "msm_kernels.cuh"
#include "msm_plan.hpp"

namespace cozk {

#if defined(__CUDACC__)
// ---- block-cooperative kernels for the shallow stages.  A single thread needs ~7 us per group addition, so the stages
// that follow level 1 are bound by DEPTH: a block of ACC_TILE threads holds one partial sum per thread in shared memory
// (structure of arrays: word k of thread t at w[k][t], conflict-free) and combines them in log2(ACC_TILE) steps.
struct ShPoints {
    uint32_t w[32][ACC_TILE];
};
__device__ __forceinline__ void sh_store(ShPoints& s, int t, const xyzz& p) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s.w[k][t] = p.X.v[k];
        s.w[8 + k][t] = p.Y.v[k];
        s.w[16 + k][t] = p.ZZ.v[k];
        s.w[24 + k][t] = p.ZZZ.v[k];
    }
}
__device__ __forceinline__ xyzz sh_load(const ShPoints& s, int t) {
    xyzz p;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        p.X.v[k] = s.w[k][t];
        p.Y.v[k] = s.w[8 + k][t];
        p.ZZ.v[k] = s.w[16 + k][t];
        p.ZZZ.v[k] = s.w[24 + k][t];
    }
    return p;
}
#endif

struct TreeSumArgs {
    const xyzz* s;
    const xyzz* w;
    xyzz* out;        // [(win*NS + id)*chunks + q]
    uint32_t G;       // entries per (window[, id]) array
    uint32_t NS;
    uint32_t chunk;   // entries per block
    uint32_t chunks;  // G / chunk
    int masked;       // 1: sources are s / w with [win*G + e] and the bit masks; 0: source is s with [(win*NS + id)*G + e]
};

void launch_merge(const MergeArgs& A, unsigned blocks, cudaStream_t st);                  // 128 threads per block
void launch_group(const GroupArgs& A, unsigned blocks, cudaStream_t st);                  // 64 threads per block
void launch_treesum(const TreeSumArgs& A, unsigned blocks, cudaStream_t st);              // ACC_TILE threads per block
void launch_finish(const FinishArgs& A, unsigned blocks, cudaStream_t st);                // 32 threads per block
void launch_rowcol(const RowColArgs& A, cudaStream_t st);                                 // A.blocks blocks of ACC_TILE threads
void launch_masksum(const MaskSumArgs& A, cudaStream_t st);                               // A.blocks blocks of ACC_TILE threads

}  // namespace cozk
