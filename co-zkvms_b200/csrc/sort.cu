// Host driver of the pair sort (sort_kernels.cuh) and its test entry point.
#include "engine.hpp"
#include "sort_kernels.cuh"

namespace cozk {

static const size_t SORT_SMEM_FULL = sizeof(SortSmem);
static const size_t SORT_SMEM_COUNT = offsetof(SortSmem, delta);
static const size_t SORTGEN_SMEM_FULL = sizeof(SortGenSmem);  // the fused kernels keep the scalars' limbs behind the sort's own arrays

// Both limits are per device: called from cozk_init for every device of the context (the current device is set).
int sort_setup_device() {
    COZK_CUDA(cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM_FULL));
    COZK_CUDA(cudaFuncSetAttribute(k_sortgen_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORTGEN_SMEM_FULL));
    COZK_CUDA(cudaFuncSetAttribute(k_sort_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM_COUNT));
    COZK_CUDA(cudaFuncSetAttribute(k_sortgen_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORTGEN_SMEM_FULL));
    return COZK_OK;
}

// Groups m pairs by key on stream st (keys must be below 2^key_bits; the order inside a group is unspecified).
//   fused == nullptr: the pairs lie in (D.keys_a, D.vals_a);
//   fused != nullptr: the pairs are those of the plain decompose layout of *fused (m = g * n * W): they are generated in
//                     the first pass and never stored unsorted.  fused->keys / vals are ignored.
// The sorted pairs end up in one of the two buffer pairs of the device; *keys_out / *vals_out say which.  The caller
// has sized keys_a / vals_a / keys_b / vals_b (set 1: keys_a2 .. vals_b2) for m entries.  after_first (optional) is recorded
// behind the first pass.  Sorts of one device are serialised by their callers (they share sort_tmp).
int sort_pairs(Device& D, cudaStream_t st, const DecomposeArgs* fused, size_t m, uint32_t key_bits, uint32_t** keys_out,
               uint32_t** vals_out, double* launches, cudaEvent_t after_first, int set) {
    uint32_t* bufk[2] = {(set ? D.keys_a2 : D.keys_a).as<uint32_t>(), (set ? D.keys_b2 : D.keys_b).as<uint32_t>()};
    uint32_t* bufv[2] = {(set ? D.vals_a2 : D.vals_a).as<uint32_t>(), (set ? D.vals_b2 : D.vals_b).as<uint32_t>()};
    const SortPlan plan = SortPlan::for_bits(key_bits, (uint32_t)D.sort_digit_bits);
    if (m == 0 || m > (size_t)0x7FFFFFFF || plan.passes > SORT_MAX_PASSES || key_bits > 31) {
        set_error("internal: group too large for the sort");
        return COZK_ERR_INVALID_ARG;
    }
    uint32_t gen_blocks = 0, gen_per = 0, fus_blocks = 0, fus_per = 0;
    sort_grid((m + SORT_TILE - 1) / SORT_TILE, &gen_blocks, &gen_per);
    size_t scalars = 0;
    if (fused) {
        scalars = decompose_total(*fused);
        sort_grid((scalars + SORT_THREADS - 1) / SORT_THREADS, &fus_blocks, &fus_per);
    }
    // histogram / cursor array (reused by the passes; the last one has an entry per key value) and two `starts` arrays
    // (first positions of this pass and of the one before: the rows of this pass's scan start there)
    const size_t max_entries = (size_t)1 << plan.key_bits;
    const size_t prev_entries = plan.passes > 1 ? (size_t)1 << (plan.key_bits - plan.shift[plan.passes - 2]) : 1;
    int rc = D.sort_tmp.ensure((2 * max_entries + prev_entries + 64) * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t* cursor = D.sort_tmp.as<uint32_t>();
    uint32_t* starts[2] = {cursor + max_entries, cursor + 2 * max_entries};
    // the last pass writes its starts into starts[0] (max_entries), every earlier pass fits prev_entries: alternate so that
    // pass p reads the starts of pass p - 1 while writing its own
    auto starts_of = [&](uint32_t p) { return (plan.passes - 1 - p) % 2 == 0 ? starts[0] : starts[1]; };

    for (uint32_t p = 0; p < plan.passes; ++p) {
        SortPass P{};
        P.shift = plan.shift[p];
        P.prev_shift = plan.prev_shift[p];
        P.cursor = cursor;
        P.keys_in = bufk[p & 1];
        P.vals_in = bufv[p & 1];
        P.keys_out = bufk[(p + 1) & 1];
        P.vals_out = bufv[(p + 1) & 1];
        const size_t entries = (size_t)1 << (plan.key_bits - P.shift);
        COZK_CUDA(cudaMemsetAsync(cursor, 0, entries * sizeof(uint32_t), st));
        auto rowscan = [&](uint32_t pass) -> int {
            const uint32_t log_row = (pass == 0 ? plan.key_bits : plan.shift[pass - 1]) - plan.shift[pass];  // digit bits of this pass
            const uint32_t rows = (uint32_t)(entries >> log_row);
            k_sort_rowscan<<<(rows + 7) / 8, 256, 0, st>>>(cursor, starts_of(pass), pass == 0 ? nullptr : starts_of(pass - 1), rows, log_row);
            COZK_CUDA(cudaGetLastError());
            *launches += 1;
            return COZK_OK;
        };
        if (p == 0 && fused) {
            P.m = scalars;
            P.nblocks = fus_blocks;
            P.tiles_per_block = fus_per;
            k_sortgen_count<<<P.nblocks, SORT_THREADS, SORTGEN_SMEM_FULL, st>>>(*fused, P);
            COZK_CUDA(cudaGetLastError());
            if ((rc = rowscan(p))) return rc;
            k_sortgen_scatter<<<P.nblocks, SORT_THREADS, SORTGEN_SMEM_FULL, st>>>(*fused, P);
            COZK_CUDA(cudaGetLastError());
        } else {
            P.m = m;
            P.nblocks = gen_blocks;
            P.tiles_per_block = gen_per;
            k_sort_count<<<P.nblocks, SORT_THREADS, SORT_SMEM_COUNT, st>>>(P);
            COZK_CUDA(cudaGetLastError());
            if ((rc = rowscan(p))) return rc;
            k_sort_scatter<<<P.nblocks, SORT_THREADS, SORT_SMEM_FULL, st>>>(P);
            COZK_CUDA(cudaGetLastError());
        }
        *launches += 2;
        if (p == 0 && after_first) COZK_CUDA(cudaEventRecord(after_first, st));
    }
    *keys_out = bufk[plan.passes & 1];
    *vals_out = bufv[plan.passes & 1];
    return COZK_OK;
}

}  // namespace cozk
