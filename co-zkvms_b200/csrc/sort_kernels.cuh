// Radix sort of the (bucket key, point) pairs of the MSM pipeline - hand-written for sm_100a, replaces the library sort.
//
// Least-significant-digit radix sort with digits of at most 8 bits.  One pass = three launches:
//   count    block b counts the digit values of ITS pairs                       -> counts[digit][block]
//   scan     one block per digit value: exclusive scan of its row over the blocks (+ the row total)
//   scatter  block b ranks its pairs (stable), sorts each tile of 8192 pairs by digit in shared memory and writes every
//            digit's run to its place: position = (pairs with a smaller digit) + (pairs with this digit in earlier
//            blocks) + (in earlier tiles of this block) + (rank in the tile).
// The FIRST pass is fused with the decompose step (k_sortgen_*): the pairs are produced from the scalars in registers
// and are born partitioned by their low digit, so the unsorted pair list is never written to or read from HBM.  The
// dominant-digit layout (compacted segments, msm_kernels.cuh) keeps its own decompose kernel and sorts with generic
// passes only.
//
// Ranking uses no shared-memory atomics: the lanes of a warp that hold the same digit find each other with r ballots
// (peers mask), the lowest of them bumps the warp's private counter, every lane's rank is the old counter value plus
// the number of peers below it.  Warps own contiguous slices of a tile, so (digit, warp, iteration, lane) order is
// the input order: the sort is stable, as every pass after the first must be.
//
// HBM traffic per pair: fused pass 8 B written; every further pass 4 B (count) + 8 B read + 8 B written.  Runs written
// per (tile, digit) are 8192 / 256 = 32 pairs = 128 B per array on average.
#pragma once
#include <cstddef>
#include <cstdint>

#include "msm_kernels.cuh"

namespace cozk {

constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;                           // pairs per thread and tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;     // 8192 pairs
constexpr uint32_t SORT_RMAX = 8;                        // digit bits per pass
constexpr int SORT_DIGITS = 1 << SORT_RMAX;
constexpr uint32_t SORT_MAX_PASSES = 4;                  // keys have at most 31 bits
constexpr uint32_t SORT_TARGET_BLOCKS = 1184;            // 148 SMs x 2 resident blocks x 4 waves

struct SortSmem {
    uint32_t wc[SORT_WARPS][SORT_DIGITS];  // per-warp digit counts; after the prefix step: offset of (digit, warp) inside the digit's run
    uint32_t lstart[SORT_DIGITS];          // start of the digit's run inside the tile
    uint32_t cur[SORT_DIGITS];             // global position of the digit's next pair (this block)
    uint32_t delta[SORT_DIGITS];           // global position minus tile position, per digit
    uint32_t wsum[SORT_WARPS];
    uint2 staged[SORT_TILE];               // the tile, sorted by digit
};

struct SortPass {
    uint32_t shift, r;        // digit = (key >> shift) & ((1 << r) - 1)
    uint32_t nblocks;         // blocks of the count / scatter grids
    uint32_t tiles_per_block; // generic passes: tiles of SORT_TILE pairs per block; fused pass: chunks of SORT_THREADS scalars
    size_t m;                 // pairs (generic passes) / scalars g * n (fused pass)
    uint32_t* counts;         // [2^r][nblocks]; after the scan: exclusive prefix along the blocks
    uint32_t* totals;         // [2^r] pairs per digit value
    const uint32_t* keys_in;
    const uint32_t* vals_in;
    uint32_t* keys_out;
    uint32_t* vals_out;
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t sort_lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// lanes of the warp that are valid and hold the same r-bit digit as the caller (garbage for invalid callers)
__device__ __forceinline__ uint32_t sort_peers(uint32_t d, bool valid, uint32_t r) {
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
    for (uint32_t b = 0; b < SORT_RMAX; ++b) {
        if (b < r) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
            peers &= bit ? bal : ~bal;
        }
    }
    return peers;
}

// count step of one warp iteration: the lowest peer adds the group's size to the warp's counter
__device__ __forceinline__ void sort_count_step(uint32_t* wc_warp, uint32_t d, bool valid, uint32_t r, uint32_t lane) {
    const uint32_t peers = sort_peers(d, valid, r);
    if (valid && lane == (uint32_t)(__ffs(peers) - 1)) wc_warp[d] += (uint32_t)__popc(peers);
    __syncwarp();
}

// rank step: as above, and every lane learns its rank among the warp's pairs with this digit so far
__device__ __forceinline__ uint32_t sort_rank_step(uint32_t* wc_warp, uint32_t d, bool valid, uint32_t r, uint32_t lane) {
    const uint32_t peers = sort_peers(d, valid, r);
    const uint32_t leader = (uint32_t)(__ffs(peers) - 1) & 31u;
    uint32_t old = 0;
    if (valid && lane == leader) {
        old = wc_warp[d];
        wc_warp[d] = old + (uint32_t)__popc(peers);
    }
    old = __shfl_sync(0xFFFFFFFFu, old, leader);
    __syncwarp();
    return old + (uint32_t)__popc(peers & sort_lanemask_lt());
}

// exclusive scan of one value per thread over the first 256 threads (8 warps); other threads pass through
__device__ __forceinline__ uint32_t sort_scan256(uint32_t x, uint32_t* wsum, uint32_t t) {
    const uint32_t lane = t & 31u, warp = t >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    if (lane == 31 && warp < 8) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t w = 0; w < warp && w < 8; ++w) base += wsum[w];
    __syncthreads();
    return base + inc - x;
}

__device__ __forceinline__ void sort_zero_counters(SortSmem& S, uint32_t t) {
    uint32_t* flat = &S.wc[0][0];
#pragma unroll
    for (int k = 0; k < SORT_WARPS * SORT_DIGITS / SORT_THREADS; ++k) flat[k * SORT_THREADS + t] = 0;
}

// Block prologue of a scatter kernel: the digit's global start (pairs with a smaller digit + pairs with this digit in
// earlier blocks) becomes the block's write cursor.
__device__ __forceinline__ void sort_init_cursors(SortSmem& S, const SortPass& P, uint32_t t) {
    const uint32_t nd = 1u << P.r;
    const uint32_t tot = (t < nd) ? P.totals[t] : 0u;
    const uint32_t off = sort_scan256(tot, S.wsum, t);
    if (t < SORT_DIGITS) S.cur[t] = (t < nd) ? off + P.counts[(size_t)t * P.nblocks + blockIdx.x] : 0u;
    __syncthreads();
}

// One tile: the thread's pairs (key[j], val[j], valid bit j of vmask) are ranked, sorted by digit in shared memory and
// written out.  Item j of lane l of warp w is element w * 512 + 32 * j + l of the tile in input order.
// The counters must be zero on entry (and the block synchronised); they are zero again - and the block synchronised -
// on return.
__device__ __forceinline__ void sort_scatter_tile(SortSmem& S, const SortPass& P, const uint32_t (&key)[SORT_ITEMS],
                                                  const uint32_t (&val)[SORT_ITEMS], uint32_t vmask, uint32_t t) {
    const uint32_t lane = t & 31u, warp = t >> 5;
    const uint32_t dmask = (1u << P.r) - 1u;
    uint32_t rank2[SORT_ITEMS / 2];  // ranks are below 8192: two per register
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const uint32_t d = (key[j] >> P.shift) & dmask;
        const uint32_t rk = sort_rank_step(S.wc[warp], d, (vmask >> j) & 1u, P.r, lane);
        if (j & 1) rank2[j / 2] |= rk << 16;
        else rank2[j / 2] = rk;
    }
    __syncthreads();
    // per digit: prefix over the warps, then the digit's start in the tile
    uint32_t total = 0;
    if (t < SORT_DIGITS) {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t c = S.wc[w][t];
            S.wc[w][t] = total;
            total += c;
        }
    }
    const uint32_t ls = sort_scan256(total, S.wsum, t);
    if (t < SORT_DIGITS) {
        S.lstart[t] = ls;
        S.delta[t] = S.cur[t] - ls;
        S.cur[t] += total;
    }
    // the tile's pair count: start + total of the last digit value
    if (t == SORT_DIGITS - 1) S.wsum[SORT_WARPS - 1] = ls + total;
    __syncthreads();
    const uint32_t tile_count = S.wsum[SORT_WARPS - 1];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        if ((vmask >> j) & 1u) {
            const uint32_t d = (key[j] >> P.shift) & dmask;
            const uint32_t rk = (rank2[j / 2] >> (16 * (j & 1))) & 0xFFFFu;
            S.staged[S.lstart[d] + S.wc[warp][d] + rk] = make_uint2(key[j], val[j]);
        }
    }
    __syncthreads();
    sort_zero_counters(S, t);
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const uint32_t idx = (uint32_t)j * SORT_THREADS + t;
        if (idx < tile_count) {
            const uint2 kv = S.staged[idx];
            const uint32_t g = S.delta[(kv.x >> P.shift) & dmask] + idx;
            P.keys_out[g] = kv.x;
            P.vals_out[g] = kv.y;
        }
    }
    __syncthreads();
}

// end of a count kernel: sum the warps' counters, one row entry per digit value
__device__ __forceinline__ void sort_store_counts(SortSmem& S, const SortPass& P, uint32_t t) {
    __syncthreads();
    if (t < (1u << P.r)) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) total += S.wc[w][t];
        P.counts[(size_t)t * P.nblocks + blockIdx.x] = total;
    }
}

// ------------------------------------------------------------------------------------------------ generic passes
__global__ void __launch_bounds__(SORT_THREADS, 2) k_sort_count(SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(sort_smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t dmask = (1u << P.r) - 1u;
    sort_zero_counters(S, t);
    __syncthreads();
    const size_t first = (size_t)blockIdx.x * P.tiles_per_block * SORT_TILE;
    for (uint32_t tile = 0; tile < P.tiles_per_block; ++tile) {
        const size_t base = first + (size_t)tile * SORT_TILE + (size_t)warp * (32 * SORT_ITEMS) + lane;
        if (first + (size_t)tile * SORT_TILE >= P.m) break;
        uint32_t key[SORT_ITEMS];
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            const size_t i = base + 32 * (size_t)j;
            key[j] = i < P.m ? P.keys_in[i] : 0u;
        }
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j)
            sort_count_step(S.wc[warp], (key[j] >> P.shift) & dmask, base + 32 * (size_t)j < P.m, P.r, lane);
    }
    sort_store_counts(S, P, t);
}

// one block per digit value: exclusive scan of the row counts[digit][0 .. nblocks) in place, row total to totals[digit]
__global__ void __launch_bounds__(256) k_sort_scan(SortPass P) {
    __shared__ uint32_t wsum[SORT_WARPS];
    __shared__ uint32_t carry_s;
    uint32_t* row = P.counts + (size_t)blockIdx.x * P.nblocks;
    const uint32_t t = threadIdx.x;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < P.nblocks; base += 256) {
        const uint32_t i = base + t;
        const uint32_t x = i < P.nblocks ? row[i] : 0u;
        const uint32_t ex = sort_scan256(x, wsum, t);
        const uint32_t carry = carry_s;
        if (i < P.nblocks) row[i] = carry + ex;
        __syncthreads();
        if (t == 255) carry_s = carry + ex + x;
        __syncthreads();
    }
    if (t == 0) P.totals[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(SORT_THREADS, 2) k_sort_scatter(SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(sort_smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    sort_zero_counters(S, t);
    sort_init_cursors(S, P, t);
    const size_t first = (size_t)blockIdx.x * P.tiles_per_block * SORT_TILE;
    for (uint32_t tile = 0; tile < P.tiles_per_block; ++tile) {
        if (first + (size_t)tile * SORT_TILE >= P.m) break;
        const size_t base = first + (size_t)tile * SORT_TILE + (size_t)warp * (32 * SORT_ITEMS) + lane;
        uint32_t key[SORT_ITEMS], val[SORT_ITEMS], vmask = 0;
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            const size_t i = base + 32 * (size_t)j;
            const bool ok = i < P.m;
            key[j] = ok ? P.keys_in[i] : 0u;
            val[j] = ok ? P.vals_in[i] : 0u;
            vmask |= (uint32_t)ok << j;
        }
        sort_scatter_tile(S, P, key, val, vmask, t);
    }
}

// ------------------------------------------------------------------------------------------------ fused first pass
// P.m = g * n scalars; block b owns scalars [b * tiles_per_block * 512, ...), one scalar per thread and chunk; the pairs
// of a chunk go through the tile machinery in groups of 16 windows.  Both kernels derive the pairs exactly as
// decompose_body does in its plain layout (make_pair).
__global__ void __launch_bounds__(SORT_THREADS, 2) k_sortgen_count(DecomposeArgs A, SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(sort_smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t dmask = (1u << P.r) - 1u;
    sort_zero_counters(S, t);
    __syncthreads();
    const size_t first = (size_t)blockIdx.x * P.tiles_per_block * SORT_THREADS;
    for (uint32_t chunk = 0; chunk < P.tiles_per_block; ++chunk) {
        const size_t cbase = first + (size_t)chunk * SORT_THREADS;
        if (cbase >= P.m) break;
        const size_t tid = cbase + t;
        const bool ok = tid < P.m;
        uint32_t v = 0;
        size_t i = 0;
        fr s = fq_zero();
        bool skip = false;
        if (ok) {
            s = decompose_load(tid, A, v, i);
            skip = A.infinity && A.infinity[i];
        }
        uint32_t carry = 0;
        for (uint32_t w = 0; w < A.W; ++w) {
            uint32_t neg, key, val;
            const uint32_t d = signed_digit(s, w, A.c, carry, neg);
            make_pair(A, v, i, w, d, neg, skip, key, val);
            sort_count_step(S.wc[warp], (key >> P.shift) & dmask, ok, P.r, lane);
        }
    }
    sort_store_counts(S, P, t);
}

__global__ void __launch_bounds__(SORT_THREADS, 2) k_sortgen_scatter(DecomposeArgs A, SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(sort_smem_raw);
    const uint32_t t = threadIdx.x;
    sort_zero_counters(S, t);
    sort_init_cursors(S, P, t);
    const size_t first = (size_t)blockIdx.x * P.tiles_per_block * SORT_THREADS;
    for (uint32_t chunk = 0; chunk < P.tiles_per_block; ++chunk) {
        const size_t cbase = first + (size_t)chunk * SORT_THREADS;
        if (cbase >= P.m) break;
        const size_t tid = cbase + t;
        const bool ok = tid < P.m;
        uint32_t v = 0;
        size_t i = 0;
        fr s = fq_zero();
        bool skip = false;
        if (ok) {
            s = decompose_load(tid, A, v, i);
            skip = A.infinity && A.infinity[i];
        }
        uint32_t carry = 0;
        for (uint32_t w0 = 0; w0 < A.W; w0 += SORT_ITEMS) {
            uint32_t key[SORT_ITEMS], val[SORT_ITEMS], vmask = 0;
#pragma unroll
            for (int j = 0; j < SORT_ITEMS; ++j) {
                key[j] = 0;
                val[j] = 0;
                if (w0 + j < A.W) {
                    uint32_t neg;
                    const uint32_t d = signed_digit(s, w0 + j, A.c, carry, neg);
                    make_pair(A, v, i, w0 + j, d, neg, skip, key[j], val[j]);
                    vmask |= (uint32_t)ok << j;
                }
            }
            sort_scatter_tile(S, P, key, val, vmask, t);
        }
    }
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------ host-side plan
struct SortPlan {
    uint32_t passes = 0;
    uint32_t shift[SORT_MAX_PASSES] = {}, r[SORT_MAX_PASSES] = {};
    // digit widths as even as possible: 16 bits -> 8 + 8, 21 -> 7 + 7 + 7
    static SortPlan for_bits(uint32_t key_bits) {
        SortPlan p;
        if (key_bits == 0) key_bits = 1;
        p.passes = (key_bits + SORT_RMAX - 1) / SORT_RMAX;
        uint32_t done = 0;
        for (uint32_t i = 0; i < p.passes; ++i) {
            const uint32_t left = p.passes - i;
            p.r[i] = (key_bits - done + left - 1) / left;
            p.shift[i] = done;
            done += p.r[i];
        }
        return p;
    }
};

// grid of a pass over `units` tiles (generic: tiles of SORT_TILE pairs; fused: chunks of SORT_THREADS scalars)
inline void sort_grid(size_t units, uint32_t* nblocks, uint32_t* per_block) {
    size_t per = (units + SORT_TARGET_BLOCKS - 1) / SORT_TARGET_BLOCKS;
    if (per == 0) per = 1;
    *per_block = (uint32_t)per;
    *nblocks = (uint32_t)((units + per - 1) / per);
    if (*nblocks == 0) *nblocks = 1;
}

}  // namespace cozk
