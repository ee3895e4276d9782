// Host driver of the pair sort (sort_kernels.cuh) and its test entry point.
#include "engine.hpp"
#include "sort_kernels.cuh"

namespace cozk {

static const size_t SORT_SMEM_FULL = sizeof(SortSmem);
static const size_t SORT_SMEM_COUNT = offsetof(SortSmem, staged);

// Both limits are per device: called from cozk_init for every device of the context (the current device is set).
int sort_setup_device() {
    COZK_CUDA(cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM_FULL));
    COZK_CUDA(cudaFuncSetAttribute(k_sortgen_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM_FULL));
    COZK_CUDA(cudaFuncSetAttribute(k_sort_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM_COUNT));
    COZK_CUDA(cudaFuncSetAttribute(k_sortgen_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM_COUNT));
    return COZK_OK;
}

// Sorts m pairs by the low key_bits bits of their keys on stream st.
//   fused == nullptr: the pairs lie in (D.keys_a, D.vals_a);
//   fused != nullptr: the pairs are those of the plain decompose layout of *fused (m = g * n * W): they are generated in
//                     the first pass and never stored unsorted.  fused->keys / vals are ignored.
// The sorted pairs end up in one of the two buffer pairs of the device; *keys_out / *vals_out say which.  The caller
// has sized keys_a / vals_a / keys_b / vals_b for m entries.  after_first (optional) is recorded behind the first pass.
int sort_pairs(Device& D, cudaStream_t st, const DecomposeArgs* fused, size_t m, uint32_t key_bits, uint32_t** keys_out,
               uint32_t** vals_out, double* launches, cudaEvent_t after_first) {
    uint32_t* bufk[2] = {D.keys_a.as<uint32_t>(), D.keys_b.as<uint32_t>()};
    uint32_t* bufv[2] = {D.vals_a.as<uint32_t>(), D.vals_b.as<uint32_t>()};
    const SortPlan plan = SortPlan::for_bits(key_bits);
    if (m == 0 || m > (size_t)0x7FFFFFFF || plan.passes > SORT_MAX_PASSES) {
        set_error("internal: group too large for the sort");
        return COZK_ERR_INVALID_ARG;
    }
    // row storage for the largest grid of this call
    uint32_t gen_blocks = 0, gen_per = 0, fus_blocks = 0, fus_per = 0;
    sort_grid((m + SORT_TILE - 1) / SORT_TILE, &gen_blocks, &gen_per);
    size_t scalars = 0;
    if (fused) {
        scalars = (size_t)fused->g * fused->n;
        sort_grid((scalars + SORT_THREADS - 1) / SORT_THREADS, &fus_blocks, &fus_per);
    }
    const size_t max_blocks = gen_blocks > fus_blocks ? gen_blocks : fus_blocks;
    int rc = D.sort_tmp.ensure((SORT_DIGITS * max_blocks + SORT_DIGITS) * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t* counts = D.sort_tmp.as<uint32_t>();
    uint32_t* totals = counts + SORT_DIGITS * max_blocks;

    for (uint32_t p = 0; p < plan.passes; ++p) {
        SortPass P{};
        P.shift = plan.shift[p];
        P.r = plan.r[p];
        P.counts = counts;
        P.totals = totals;
        P.keys_in = bufk[p & 1];
        P.vals_in = bufv[p & 1];
        P.keys_out = bufk[(p + 1) & 1];
        P.vals_out = bufv[(p + 1) & 1];
        if (p == 0 && fused) {
            P.m = scalars;
            P.nblocks = fus_blocks;
            P.tiles_per_block = fus_per;
            k_sortgen_count<<<P.nblocks, SORT_THREADS, SORT_SMEM_COUNT, st>>>(*fused, P);
            COZK_CUDA(cudaGetLastError());
            k_sort_scan<<<1u << P.r, 256, 0, st>>>(P);
            COZK_CUDA(cudaGetLastError());
            k_sortgen_scatter<<<P.nblocks, SORT_THREADS, SORT_SMEM_FULL, st>>>(*fused, P);
            COZK_CUDA(cudaGetLastError());
        } else {
            P.m = m;
            P.nblocks = gen_blocks;
            P.tiles_per_block = gen_per;
            k_sort_count<<<P.nblocks, SORT_THREADS, SORT_SMEM_COUNT, st>>>(P);
            COZK_CUDA(cudaGetLastError());
            k_sort_scan<<<1u << P.r, 256, 0, st>>>(P);
            COZK_CUDA(cudaGetLastError());
            k_sort_scatter<<<P.nblocks, SORT_THREADS, SORT_SMEM_FULL, st>>>(P);
            COZK_CUDA(cudaGetLastError());
        }
        *launches += 3;
        if (p == 0 && after_first) COZK_CUDA(cudaEventRecord(after_first, st));
    }
    *keys_out = bufk[plan.passes & 1];
    *vals_out = bufv[plan.passes & 1];
    return COZK_OK;
}

}  // namespace cozk
