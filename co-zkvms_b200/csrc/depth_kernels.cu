// The stages of the MSM pipeline behind the accumulate stage: the bucket merge of the streamed mode, the bucket reduce
// and the on-device finish.  Thread bodies: msm_kernels.cuh.
//
// They are bound by DEPTH, not by throughput: a few thousand threads each walk a chain of a few dozen group additions.
// With the field operations expanded inline a group addition is ~35 KB of straight-line code, and a scheduler with one
// or two resident warps waits on instruction fetch for much of it; the accumulate kernels hide that behind their warps
// and want the inline form (level 1: 2.62 against 2.80 ms at 2^20 with calls), these kernels do not.  This translation
// unit is therefore compiled with COZK_FIELD_CALLS: fq_mul / fq_sqr / fq_mul2 become real calls (register ABI, no
// stack), the kernels shrink to a few KB and stay in the instruction cache.  Measured: bucket reduce 0.24 -> 0.21 ms at
// 2^16 .. 2^20, 0.20 -> 0.15 ms at 2^12; whole MSM 2^12 0.615 -> 0.568 ms, 2^16 0.750 -> 0.718 ms; PST13 opening at
// nv = 18 3.90 -> 3.73 ms.
// (The macro only changes code under __CUDA_ARCH__: the host bodies of the inline field functions are the same in every
// translation unit, and device code is compiled per translation unit - no relocatable device code, nothing is linked
// across units - so the two device forms of fq_mul never meet.)
#define COZK_FIELD_CALLS 1
#include "depth_kernels.hpp"

#include "msm_plan.hpp"

namespace cozk {

__global__ void __launch_bounds__(128, 3) k_merge(MergeArgs A) {
    merge_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}
__global__ void __launch_bounds__(128, 3) k_group(GroupArgs A) {
    group_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}

// Bucket-reduce sums: block (q, id, win) adds the groups of chunk q that belong to sum `id` of window `win`
// (id 0: all W_g, id 1: all S_g, id 2+j: the S_g whose index has bit j set; plain = no masks, one source array).
// Output contract = bitsum_body with f = chunk (masked) / plainsum_body with f = chunk (plain).
__global__ void __launch_bounds__(ACC_TILE) k_treesum(TreeSumArgs A) {
    __shared__ ShPoints sp;
    const int t = threadIdx.x;
    // blockIdx.x = (win * NS + id) * chunks + q   (one-dimensional: windows can exceed the 65,535 limit of grid.y/z)
    const uint32_t q = blockIdx.x % A.chunks;
    const size_t rest = blockIdx.x / A.chunks;
    const uint32_t id = (uint32_t)(rest % A.NS);
    const size_t win = rest / A.NS;
    const xyzz* src = A.masked ? ((id == 0 ? A.w : A.s) + win * A.G) : (A.s + (win * A.NS + id) * (size_t)A.G);
    xyzz acc = xyzz_identity();
    for (uint32_t e = q * A.chunk + t; e < (q + 1) * A.chunk; e += ACC_TILE) {
        if (A.masked && id >= 2 && !((e >> (id - 2)) & 1u)) continue;
        acc = xyzz_add(acc, load_xyzz(&src[e]));
    }
    sh_store(sp, t, acc);
    __syncthreads();
    for (int d = ACC_TILE / 2; d > 0; d >>= 1) {
        if (t < d) {
            acc = xyzz_add(acc, sh_load(sp, t + d));
            sh_store(sp, t, acc);
        }
        __syncthreads();
    }
    if (t == 0) store_xyzz(&A.out[(win * A.NS + id) * (size_t)A.chunks + q], acc);
}
// block-wide sum of one value per thread (tree in shared memory); valid in thread 0
__device__ __forceinline__ xyzz block_tree_sum(ShPoints& sp, xyzz acc) {
    const int t = threadIdx.x;
    sh_store(sp, t, acc);
    __syncthreads();
    for (int d = ACC_TILE / 2; d > 0; d >>= 1) {
        if (t < d) {
            acc = xyzz_add(acc, sh_load(sp, t + d));
            sh_store(sp, t, acc);
        }
        __syncthreads();
    }
    return acc;
}
// Row / column form of the bucket reduce (MsmPlan::reduce_2d; output contracts: rowcol_body, masksum_body).
__global__ void __launch_bounds__(ACC_TILE) k_rowcol(RowColArgs A) {
    __shared__ ShPoints sp;
    const size_t rows = (size_t)1 << A.hi_bits, cols = (size_t)1 << A.lo_bits;
    const size_t blk = blockIdx.x, win = blk / (rows + cols), i = blk % (rows + cols);
    const xyzz* b = A.buckets + win * rows * cols;
    xyzz acc = xyzz_identity();
    if (i < rows) {
        for (size_t e = threadIdx.x; e < cols; e += ACC_TILE) acc = xyzz_add(acc, load_xyzz(&b[i * cols + e]));
    } else {
        for (size_t e = threadIdx.x; e < rows; e += ACC_TILE) acc = xyzz_add(acc, load_xyzz(&b[e * cols + (i - rows)]));
    }
    acc = block_tree_sum(sp, acc);
    if (threadIdx.x == 0) store_xyzz(&A.rc[blk], acc);
}
__global__ void __launch_bounds__(ACC_TILE) k_masksum(MaskSumArgs A) {
    __shared__ ShPoints sp;
    const size_t rows = (size_t)1 << A.hi_bits, cols = (size_t)1 << A.lo_bits;
    const size_t blk = blockIdx.x, win = blk / A.NS;
    const uint32_t id = (uint32_t)(blk % A.NS);
    const xyzz* r = A.rc + win * (rows + cols);
    const xyzz* c = r + rows;
    xyzz acc = xyzz_identity();
    if (id == 1) {
        for (size_t e = threadIdx.x; e < cols; e += ACC_TILE) acc = xyzz_add(acc, load_xyzz(&c[e]));
    } else if (id >= 2 && id - 2 < A.lo_bits) {
        for (size_t e = threadIdx.x; e < cols; e += ACC_TILE)
            if ((e >> (id - 2)) & 1u) acc = xyzz_add(acc, load_xyzz(&c[e]));
    } else if (id >= 2) {
        for (size_t e = threadIdx.x; e < rows; e += ACC_TILE)
            if ((e >> (id - 2 - A.lo_bits)) & 1u) acc = xyzz_add(acc, load_xyzz(&r[e]));
    }
    acc = block_tree_sum(sp, acc);
    if (threadIdx.x == 0) store_xyzz(&A.out[blk], acc);
}
__global__ void __launch_bounds__(32) k_finish(FinishArgs A) {
    finish_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}


void launch_merge(const MergeArgs& A, unsigned blocks, cudaStream_t st) { k_merge<<<blocks, 128, 0, st>>>(A); }
void launch_group(const GroupArgs& A, unsigned blocks, cudaStream_t st) { k_group<<<blocks, 64, 0, st>>>(A); }
void launch_treesum(const TreeSumArgs& A, unsigned blocks, cudaStream_t st) { k_treesum<<<blocks, ACC_TILE, 0, st>>>(A); }
void launch_finish(const FinishArgs& A, unsigned blocks, cudaStream_t st) { k_finish<<<blocks, 32, 0, st>>>(A); }
void launch_rowcol(const RowColArgs& A, cudaStream_t st) { k_rowcol<<<(unsigned)A.blocks, ACC_TILE, 0, st>>>(A); }
void launch_masksum(const MaskSumArgs& A, cudaStream_t st) { k_masksum<<<(unsigned)A.blocks, ACC_TILE, 0, st>>>(A); }

}  // namespace cozk
