// Radix sort of the (bucket key, point) pairs of the MSM pipeline - hand-written for sm_100a, replaces the library sort.
//
// Most-significant-digit radix sort, digits of at most 8 bits, top digit first.  Pass p groups the pairs by
// prefix_p = key >> shift_p (shift_1 > shift_2 > ... > 0 = the whole key); its input is already grouped by the shorter
// prefix of the pass before.  Nothing has to be stable top-down - the order of the points inside a bucket does not matter
// to a sum - so ranks come from plain shared-memory atomics (one ATOMS per pair) instead of ballot matching.
// One pass = three steps:
//   count    histogram of prefix_p over all pairs (tile-local shared-memory counters, flushed with global reductions)
//   scan     exclusive scan of the histogram -> first position of every prefix value
//   scatter  a block takes a tile of 8192 pairs, ranks them per prefix value in shared memory, reserves room for each of
//            its runs with ONE global atomic per (tile, prefix value), sorts the tile in shared memory and writes every
//            run to its place (coalesced: a run is 8192 / 256 = 32 pairs = 128 B per array on average).
// A tile's pairs share the prefix of the pass before up to a few neighbouring values, so their prefix_p values fall into
// a window of a few hundred consecutive values: shared-memory counters cover SORT_RANGE values from the tile's smallest
// possible one; a pair outside the window (only when partitions are far smaller than a tile: tiny or very skewed inputs)
// takes the slow path - a global atomic and a scattered store of its own.
// The FIRST pass is fused with the decompose step (k_sortgen_*): the pairs are produced from the scalars in registers
// and are born partitioned by their top digit, so the unsorted pair list is never written to or read from HBM.  The
// dominant-digit layout (compacted segments, msm_kernels.cuh) keeps its own decompose kernel and sorts with generic
// passes only.
// Constant share vectors (co-jolt, dense_mlpoly.rs:567-581) send whole warps to ONE counter: a warp whose lanes all hold the
// same prefix value takes its 32 ranks with a single atomic.
//
// HBM traffic per pair: fused pass 8 B written; every further pass 4 B (count) + 8 B read + 8 B written.
#pragma once
#include <cstddef>
#include <cstdint>

#include "msm_kernels.cuh"

namespace cozk {

#ifndef COZK_SORT_THREADS       // build-time tuning knobs (tools/gpu_sortbench.sh builds variants side by side)
#define COZK_SORT_THREADS 512
#endif
#ifndef COZK_SORT_MINBLOCKS
#define COZK_SORT_MINBLOCKS 2
#endif
constexpr int SORT_THREADS = COZK_SORT_THREADS;
#ifndef COZK_SORT_ITEMS
#define COZK_SORT_ITEMS 16
#endif
constexpr int SORT_ITEMS = COZK_SORT_ITEMS;              // pairs per thread and tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;     // 8192 pairs
constexpr uint32_t SORT_RMAX = 8;                        // digit bits per pass (default; option "sort_digit_bits")
constexpr uint32_t SORT_RMAX_LIMIT = 11;                 // 2^11 = SORT_RANGE prefix values: what a tile's shared-memory counters cover
constexpr uint32_t SORT_RANGE = 2048;                    // prefix values a tile ranks through shared memory
constexpr int SORT_PER_THREAD = SORT_RANGE / SORT_THREADS;
constexpr uint32_t SORT_MAX_PASSES = 4;                  // keys have at most 31 bits
constexpr uint32_t SORT_TARGET_BLOCKS = 592 * COZK_SORT_MINBLOCKS;  // 148 SMs x resident blocks x 4 waves

struct SortSmem {
    uint32_t cnt[SORT_RANGE];     // pairs per prefix value (relative to the tile's base); then: start of its run in the tile
    uint32_t delta[SORT_RANGE];   // global position minus tile position, per prefix value
    uint32_t wsum[32];
    uint2 staged[SORT_TILE];      // the tile, grouped by prefix value
};

struct SortPass {
    uint32_t shift;           // this pass groups by key >> shift
    uint32_t prev_shift;      // the input is grouped by key >> prev_shift (32: not at all - the first pass)
    uint32_t nblocks;         // blocks of the count / scatter grids
    uint32_t tiles_per_block; // generic passes: tiles of SORT_TILE pairs per block; fused pass: chunks of SORT_THREADS scalars
    size_t m;                 // pairs (generic passes) / scalars g * n (fused pass)
    uint32_t* cursor;         // [1 << (key_bits - shift)]: histogram (count), first positions (scan), next free position (scatter)
    const uint32_t* keys_in;
    const uint32_t* vals_in;
    uint32_t* keys_out;
    uint32_t* vals_out;
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t sort_lane() { return threadIdx.x & 31u; }

// smallest prefix value a tile that starts with `first_key` can hold
__device__ __forceinline__ uint32_t sort_tile_base(const SortPass& P, uint32_t first_key) {
    if (P.prev_shift >= 32) return 0;
    return (first_key >> P.prev_shift) << (P.prev_shift - P.shift);
}

// Fused pass: a warp step holds one window of 32 neighbouring scalars - for a constant share vector 32 equal prefix
// values.  A warp whose lanes all take part with the same value shares ONE shared-memory atomic.  All 32 lanes call.
__device__ __forceinline__ uint32_t sort_take_rank_warp(uint32_t* cnt, uint32_t rel, bool in) {
    const uint32_t rel0 = __shfl_sync(0xFFFFFFFFu, rel, 0);
    if (__all_sync(0xFFFFFFFFu, in && rel == rel0)) {
        uint32_t base = 0;
        if (sort_lane() == 0) base = atomicAdd(&cnt[rel], 32u);
        return __shfl_sync(0xFFFFFFFFu, base, 0) + sort_lane();
    }
    return in ? atomicAdd(&cnt[rel], 1u) : 0u;
}
__device__ __forceinline__ void sort_count_warp(uint32_t* cnt, uint32_t rel, bool in) {
    const uint32_t rel0 = __shfl_sync(0xFFFFFFFFu, rel, 0);
    if (__all_sync(0xFFFFFFFFu, in && rel == rel0)) {
        if (sort_lane() == 0) atomicAdd(&cnt[rel], 32u);
    } else if (in) {
        atomicAdd(&cnt[rel], 1u);
    }
}

// exclusive scan of one value per thread over the block (blockDim.x <= 1024); total of the block to *total_out
__device__ __forceinline__ uint32_t sort_block_scan(uint32_t x, uint32_t* wsum, uint32_t* total_out) {
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = 0, total = 0;
    for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t sw = wsum[w];
        if (w < warp) base += sw;
        total += sw;
    }
    __syncthreads();
    if (total_out) *total_out = total;
    return base + inc - x;
}

__device__ __forceinline__ void sort_zero_counters(SortSmem& S) {
#pragma unroll
    for (int k = 0; k < SORT_PER_THREAD; ++k) S.cnt[k * SORT_THREADS + threadIdx.x] = 0;
}

// counters of a tile -> global histogram
__device__ __forceinline__ void sort_flush_counts(SortSmem& S, const SortPass& P, uint32_t base) {
#pragma unroll
    for (int k = 0; k < SORT_PER_THREAD; ++k) {
        const uint32_t c = k * SORT_THREADS + threadIdx.x, v = S.cnt[c];
        if (v) atomicAdd(&P.cursor[base + c], v);
    }
}

// Second half of a tile, after every pair has its rank: runs inside the tile (exclusive scan of the counters), one global
// atomic per run for its place in the output; returns the number of pairs that went through shared memory.
// On return cnt[e] = start of run e inside the tile, delta[e] = global position minus tile position; block synchronised.
__device__ __forceinline__ uint32_t sort_place_runs(SortSmem& S, const SortPass& P, uint32_t base) {
    const uint32_t t = threadIdx.x;
    uint32_t c[SORT_PER_THREAD], sum = 0;
#pragma unroll
    for (int k = 0; k < SORT_PER_THREAD; ++k) {
        c[k] = S.cnt[t * SORT_PER_THREAD + k];
        sum += c[k];
    }
    uint32_t tile_count;
    uint32_t start = sort_block_scan(sum, S.wsum, &tile_count);
#pragma unroll
    for (int k = 0; k < SORT_PER_THREAD; ++k) {
        const uint32_t e = t * SORT_PER_THREAD + k;
        if (c[k]) S.delta[e] = atomicAdd(&P.cursor[base + e], c[k]) - start;
        S.cnt[e] = start;
        start += c[k];
    }
    __syncthreads();
    return tile_count;
}

// Last step of a tile: the staged pairs, grouped by prefix value, go to their runs (coalesced).  Synchronises at the end.
__device__ __forceinline__ void sort_write_tile(SortSmem& S, const SortPass& P, uint32_t base, uint32_t tile_count) {
    const uint32_t t = threadIdx.x;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const uint32_t idx = (uint32_t)j * SORT_THREADS + t;
        if (idx < tile_count) {
            const uint2 kv = S.staged[idx];
            const uint32_t g = S.delta[(kv.x >> P.shift) - base] + idx;
            P.keys_out[g] = kv.x;
            P.vals_out[g] = kv.y;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ generic passes
__global__ void __launch_bounds__(SORT_THREADS, COZK_SORT_MINBLOCKS) k_sort_count(SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(sort_smem_raw);
    const uint32_t t = threadIdx.x, m = (uint32_t)P.m;
    const uint32_t first = blockIdx.x * P.tiles_per_block * SORT_TILE;
    for (uint32_t tile = 0; tile < P.tiles_per_block; ++tile) {
        const uint32_t tbase = first + tile * SORT_TILE;
        if (tbase >= m) break;
        sort_zero_counters(S);
        const uint32_t base = sort_tile_base(P, P.keys_in[tbase]);
        const uint32_t* kin = P.keys_in + tbase + t;
        const uint32_t left = m - tbase;  // pairs from the tile's start to the end of the list
        // items j < nvalid of this thread exist (item j is pair j * SORT_THREADS + t of the tile)
        const int nvalid = left > t ? (int)((left - t + SORT_THREADS - 1) / SORT_THREADS) : 0;
        uint32_t key[SORT_ITEMS];
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) key[j] = j < nvalid ? kin[j * SORT_THREADS] : 0u;
        __syncthreads();
        uint32_t far = 0;
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            const uint32_t rel = (key[j] >> P.shift) - base;
            if (j < nvalid) {
                if (rel < SORT_RANGE) atomicAdd(&S.cnt[rel], 1u);
                else far |= 1u << j;
            }
        }
        if (far) {
#pragma unroll
            for (int j = 0; j < SORT_ITEMS; ++j)
                if ((far >> j) & 1u) atomicAdd(&P.cursor[key[j] >> P.shift], 1u);
        }
        __syncthreads();
        sort_flush_counts(S, P, base);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS, COZK_SORT_MINBLOCKS) k_sort_scatter(SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(sort_smem_raw);
    const uint32_t t = threadIdx.x, m = (uint32_t)P.m;
    const uint32_t first = blockIdx.x * P.tiles_per_block * SORT_TILE;
    for (uint32_t tile = 0; tile < P.tiles_per_block; ++tile) {
        const uint32_t tbase = first + tile * SORT_TILE;
        if (tbase >= m) break;
        sort_zero_counters(S);
        const uint32_t base = sort_tile_base(P, P.keys_in[tbase]);
        const uint32_t* kin = P.keys_in + tbase + t;
        const uint32_t* vin = P.vals_in + tbase + t;
        const uint32_t left = m - tbase;
        const int nvalid = left > t ? (int)((left - t + SORT_THREADS - 1) / SORT_THREADS) : 0;
        uint32_t key[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) key[j] = j < nvalid ? kin[j * SORT_THREADS] : 0u;
        __syncthreads();
        uint32_t far = 0;  // bit j: item j does not go through shared memory (outside the window, or past the end of the list)
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            const uint32_t rel = (key[j] >> P.shift) - base;
            rank[j] = 0;
            if (j < nvalid && rel < SORT_RANGE) rank[j] = atomicAdd(&S.cnt[rel], 1u);
            else far |= 1u << j;
        }
        if (far) {  // far outside the tile's window (tiny partitions): a place of its own for each such pair
#pragma unroll
            for (int j = 0; j < SORT_ITEMS; ++j) {
                if (((far >> j) & 1u) && j < nvalid) {
                    const uint32_t pos = atomicAdd(&P.cursor[key[j] >> P.shift], 1u);
                    P.keys_out[pos] = key[j];
                    P.vals_out[pos] = vin[j * SORT_THREADS];
                }
            }
        }
        __syncthreads();
        const uint32_t tile_count = sort_place_runs(S, P, base);
        // the vals are only needed now: loaded here, they cost no registers while the keys are ranked
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            if (!((far >> j) & 1u)) {
                const uint32_t rel = (key[j] >> P.shift) - base;
                S.staged[S.cnt[rel] + rank[j]] = make_uint2(key[j], vin[j * SORT_THREADS]);
            }
        }
        __syncthreads();
        sort_write_tile(S, P, base, tile_count);
    }
}

// ------------------------------------------------------------------------------------------------ scan
// Histogram of pass p -> first positions.  The prefix values of pass p that share one prefix value of the pass before
// form a row of 2^r <= 256 consecutive entries, and that row's pairs start where the earlier pass put its partition:
//     start[row * 2^r + e] = row_start[row] + sum_{e' < e} hist[row * 2^r + e']
// so the scan of a whole histogram is rows-many independent scans of at most 256 entries: one warp per row, no carry
// between rows, ONE launch per pass (row_start = the `starts` array the earlier pass left; the first pass has one row
// that starts at 0).  Writes both `cursor` (consumed by the scatter step) and `starts` (kept for the next pass).
__global__ void __launch_bounds__(256) k_sort_rowscan(uint32_t* cursor, uint32_t* starts, const uint32_t* row_start, uint32_t rows,
                                                      uint32_t log_row) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint32_t len = 1u << log_row;                      // entries per row (up to 2^SORT_RMAX_LIMIT)
    const uint32_t piece = len < 256 ? len : 256;            // entries the warp scans at a time
    const uint32_t per = piece > 32 ? piece >> 5 : 1;        // consecutive entries per lane
    const uint32_t lanes = piece > 32 ? 32 : piece;          // lanes that hold entries
    uint32_t* c = cursor + ((size_t)row << log_row);
    uint32_t* st = starts + ((size_t)row << log_row);
    uint32_t carry = row_start ? row_start[row] : 0u;
    for (uint32_t base = 0; base < len; base += piece) {
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = ((uint32_t)k < per && lane < lanes) ? c[base + lane * per + k] : 0u;
            sum += v[k];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        uint32_t run = carry + inc - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if ((uint32_t)k < per && lane < lanes) {
                c[base + lane * per + k] = run;
                st[base + lane * per + k] = run;
            }
            run += v[k];
        }
        carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
}

// ------------------------------------------------------------------------------------------------ fused first pass
// P.m = g * n scalars; block b owns scalars [b * tiles_per_block * 512, ...), one scalar per thread and chunk; the pairs
// of a chunk go through the tile machinery in groups of 16 windows.  Both kernels derive the pairs exactly as
// decompose_body does in its plain layout (make_pair).  First pass: the window of prefix values starts at 0 and covers
// all 2^r <= 256 of them.
//
// The canonical scalar lives in SHARED memory, limb-major (limb l of thread t at sv[l * 512 + t]: conflict-free): a window's
// bits are two LDS at a run-time limb index - indexing a register array at run time would send the scalar to local
// memory instead.  Limb 8 is zero (the top window may reach past bit 255).
struct SortGenSmem {
    SortSmem sort;
    uint32_t sv[9 * SORT_THREADS];
};

// digit of window w (signed, with the running carry) -> (key, val) of the pair; mirrors signed_digit + make_pair
__device__ __forceinline__ void sortgen_pair(const DecomposeArgs& A, const uint32_t* sv_t, uint32_t w, uint32_t& carry, uint32_t key0,
                                             uint32_t point0, uint32_t key_step, uint32_t point_step, bool skip, uint32_t& key, uint32_t& val) {
    const uint32_t off = w * A.c, limb = off >> 5, sh = off & 31u;
    const uint32_t lo = sv_t[limb * SORT_THREADS], hi = limb < 8 ? sv_t[(limb + 1) * SORT_THREADS] : 0u;
    uint32_t d = (__funnelshift_r(lo, hi, sh) & ((1u << A.c) - 1u)) + carry;
    const uint32_t B = 1u << (A.c - 1);
    uint32_t neg = 0;
    carry = 0;
    if (d > B) {
        d = (1u << A.c) - d;
        neg = VAL_NEG;
        carry = 1;
    }
    const bool zero = (d == 0) || skip;
    key = key0 + w * key_step + (zero ? 0u : d - 1);
    val = zero ? VAL_SKIP : ((point0 + w * point_step) | neg);
}

// loads scalar `tid` (if it exists), leaves its canonical limbs in shared memory, returns the per-thread pair constants
__device__ __forceinline__ void sortgen_load(const DecomposeArgs& A, uint32_t* sv_t, uint32_t tid, bool ok, uint32_t& key0, uint32_t& point0,
                                             bool& skip) {
    uint32_t v = 0;
    size_t i = 0;
    fr s = fq_zero();
    skip = false;
    if (ok) {
        s = decompose_load(tid, A, v, i);
        skip = A.infinity && A.infinity[A.rag_base ? A.rag_base[v] + i : i];
    }
#pragma unroll
    for (int l = 0; l < 8; ++l) sv_t[l * SORT_THREADS] = s.v[l];
    sv_t[8 * SORT_THREADS] = 0;
    key0 = v * A.bucket_windows * (1u << (A.c - 1));
    point0 = (uint32_t)(decompose_base(A, v) + i);
}

__global__ void __launch_bounds__(SORT_THREADS, COZK_SORT_MINBLOCKS) k_sortgen_count(DecomposeArgs A, SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortGenSmem& G = *reinterpret_cast<SortGenSmem*>(sort_smem_raw);
    SortSmem& S = G.sort;
    const uint32_t t = threadIdx.x, m = (uint32_t)P.m;
    uint32_t* sv_t = G.sv + t;
    const uint32_t key_step = A.table_stride ? 0u : (1u << (A.c - 1)), point_step = (uint32_t)A.table_stride;
    sort_zero_counters(S);
    __syncthreads();
    const uint32_t first = blockIdx.x * P.tiles_per_block * SORT_THREADS;
    for (uint32_t chunk = 0; chunk < P.tiles_per_block; ++chunk) {
        const uint32_t cbase = first + chunk * SORT_THREADS;
        if (cbase >= m) break;
        const bool ok = cbase + t < m;
        uint32_t key0, point0;
        bool skip;
        sortgen_load(A, sv_t, cbase + t, ok, key0, point0, skip);
        uint32_t carry = 0;
        for (uint32_t w = 0; w < A.W; ++w) {
            uint32_t key, val;
            sortgen_pair(A, sv_t, w, carry, key0, point0, key_step, point_step, skip, key, val);
            sort_count_warp(S.cnt, ok ? key >> P.shift : 0u, ok);
        }
    }
    __syncthreads();
    sort_flush_counts(S, P, 0);
}

__global__ void __launch_bounds__(SORT_THREADS, COZK_SORT_MINBLOCKS) k_sortgen_scatter(DecomposeArgs A, SortPass P) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortGenSmem& G = *reinterpret_cast<SortGenSmem*>(sort_smem_raw);
    SortSmem& S = G.sort;
    const uint32_t t = threadIdx.x, m = (uint32_t)P.m;
    uint32_t* sv_t = G.sv + t;
    const uint32_t key_step = A.table_stride ? 0u : (1u << (A.c - 1)), point_step = (uint32_t)A.table_stride;
    const uint32_t first = blockIdx.x * P.tiles_per_block * SORT_THREADS;
    for (uint32_t chunk = 0; chunk < P.tiles_per_block; ++chunk) {
        const uint32_t cbase = first + chunk * SORT_THREADS;
        if (cbase >= m) break;
        const bool ok = cbase + t < m;
        uint32_t key0, point0;
        bool skip;
        sortgen_load(A, sv_t, cbase + t, ok, key0, point0, skip);
        uint32_t carry = 0;
        for (uint32_t w0 = 0; w0 < A.W; w0 += SORT_ITEMS) {
            sort_zero_counters(S);
            __syncthreads();
            uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
            for (int j = 0; j < SORT_ITEMS; ++j) {
                key[j] = 0;
                val[j] = 0;
                rank[j] = 0;
                if (w0 + j < A.W) {  // uniform over the block
                    sortgen_pair(A, sv_t, w0 + j, carry, key0, point0, key_step, point_step, skip, key[j], val[j]);
                    rank[j] = sort_take_rank_warp(S.cnt, ok ? key[j] >> P.shift : 0u, ok);
                }
            }
            __syncthreads();
            const uint32_t tile_count = sort_place_runs(S, P, 0);
            if (ok) {
#pragma unroll
                for (int j = 0; j < SORT_ITEMS; ++j)
                    if (w0 + j < A.W) S.staged[S.cnt[key[j] >> P.shift] + rank[j]] = make_uint2(key[j], val[j]);
            }
            __syncthreads();
            sort_write_tile(S, P, 0, tile_count);
        }
    }
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------ host-side plan
struct SortPlan {
    uint32_t passes = 0;
    uint32_t shift[SORT_MAX_PASSES] = {}, prev_shift[SORT_MAX_PASSES] = {};
    uint32_t key_bits = 0;
    // digit widths as even as possible, top digit first: 16 bits -> 8 + 8, 21 -> 7 + 7 + 7
    static SortPlan for_bits(uint32_t key_bits, uint32_t rmax = SORT_RMAX) {
        SortPlan p;
        if (key_bits == 0) key_bits = 1;
        if (rmax < 7) rmax = 7;  // 4 passes must cover the 25 key bits of the largest group
        if (rmax > SORT_RMAX_LIMIT) rmax = SORT_RMAX_LIMIT;
        p.key_bits = key_bits;
        p.passes = (key_bits + rmax - 1) / rmax;
        uint32_t left_bits = key_bits;
        for (uint32_t i = 0; i < p.passes && i < SORT_MAX_PASSES; ++i) {
            const uint32_t left = p.passes - i;
            const uint32_t r = (left_bits + left - 1) / left;
            p.prev_shift[i] = i == 0 ? 32 : p.shift[i - 1];
            left_bits -= r;
            p.shift[i] = left_bits;
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
