// libcozk_msm.so - the engine behind include/cozk_msm.h: kernels (thin __global__ wrappers around the thread bodies
// of msm_kernels.cuh), the per-device pipeline driver, multi-device sharding and the C ABI.
//
// Reference boundary this replaces: jolt_core::msm::{msm_field_elements, batch_msm} as called from
// co-jolt/src/poly/commitment/pst13.rs:286, :319, :461 and ark_ec::VariableBaseMSM::msm_bigint as called from
// co-noir-spartan/co-spartan/src/worker.rs:585, :804.  No CPU MSM path exists in this library.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>

#include "depth_kernels.hpp"
#include "engine.hpp"
#include "msm_kernels.cuh"
#include "msm_plan.hpp"

namespace cozk {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

Device::~Device() {
    cudaSetDevice(id);
    DevBuf* bufs[] = {&scalars[0], &scalars[1], &vec_ptrs, &keys_a, &vals_a, &keys_b, &vals_b, &sort_tmp, &buckets,
                      &keys_a2, &vals_a2, &keys_b2, &vals_b2, &aff[0], &aff[1],
                      &pk[0], &pk[1], &pk[2], &pk[3], &pp[0], &pp[1], &pp[2], &pp[3], &rs[0], &rs[1], &rw[0], &rw[1], &out, &flush,
                      &buckets2[0], &buckets2[1],
                      &dom_cand, &dom_counts, &dom_mode, &dom_off, &dom_len, &dom_cursor, &rag,
                      &open_in, &open_r[0], &open_r[1], &open_q, &open_qs};
    for (DevBuf* b : bufs) b->release();
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : copy_done) if (e) cudaEventDestroy(e);
    for (auto& e : chunk_ev) if (e) cudaEventDestroy(e);
    if (aux_stream) cudaStreamDestroy(aux_stream);
    if (prep_stream) cudaStreamDestroy(prep_stream);
    if (l1b_stream) cudaStreamDestroy(l1b_stream);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
}

// ------------------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) k_decompose(DecomposeArgs A) {
    decompose_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}
__global__ void __launch_bounds__(32) k_dom_cand(DomArgs A) { dom_cand_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
// Output contract = dom_count_body over all (vector, scalar) pairs.  Block-cooperative: blockIdx.y is the vector, the
// block's threads stride over its scalars, hits are counted per warp (ballot) into shared counters and flushed with one
// global atomic per window and block - a constant vector sends EVERY scalar of a window to the same counter.
__global__ void __launch_bounds__(256) k_dom_count(DomArgs A) {
    __shared__ uint32_t sc[256];  // [0, W): candidate hits, [128, 128 + W): zero digits (W <= 128)
    const uint32_t v = blockIdx.y, W = A.D.W;
    const size_t n = A.D.n, cn = A.count_n ? A.count_n : n;
    sc[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    for (size_t base = (size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < cn; base += (size_t)gridDim.x * blockDim.x) {
        size_t i = base + lane;
        const bool valid = i < cn;
        if (cn < n && i >= cn / 2) i += n - cn;
        fr s = fq_zero();
        if (valid) {
            uint32_t vv;
            size_t ii;
            s = decompose_load((size_t)v * n + i, A.D, vv, ii);
        }
        uint32_t carry = 0;
        for (uint32_t w = 0; w < W; ++w) {
            uint32_t neg;
            const uint32_t d = signed_digit(s, w, A.D.c, carry, neg);
            const int32_t sd = neg ? -(int32_t)d : (int32_t)d;
            const unsigned hc = __ballot_sync(0xFFFFFFFFu, valid && sd == A.cand[v * W + w]);
            const unsigned hz = __ballot_sync(0xFFFFFFFFu, valid && d == 0);
            if (lane == 0) {
                if (hc) atomicAdd(&sc[w], (uint32_t)__popc(hc));
                if (hz) atomicAdd(&sc[128 + w], (uint32_t)__popc(hz));
            }
        }
    }
    __syncthreads();
    const uint32_t t = threadIdx.x;
    if (t < W && sc[t]) atomicAdd(&A.count_cand[v * W + t], sc[t]);
    if (t >= 128 && t - 128 < W && sc[t]) atomicAdd(&A.count_zero[v * W + t - 128], sc[t]);
}
// Level 1 is bound by the integer-multiply pipe and by how many warps the register file holds (117 registers per thread:
// 16 warps per SM as 4 blocks of 128).  COZK_ACC_BLOCK / COZK_ACC_MAXREG: build-time knobs for other shapes
// (tools/gpu_variants.sh measures them side by side; __launch_bounds__ alone snaps from 128 straight to 96 registers).
#ifndef COZK_ACC_BLOCK
#define COZK_ACC_BLOCK 128
#endif
template <bool LEVEL1>
__global__ void __launch_bounds__(128, 4) k_accumulate(AccumulateArgs A) {
    accumulate_body<LEVEL1>((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}
#ifdef COZK_ACC_MAXREG
__global__ void __maxnreg__(COZK_ACC_MAXREG) k_accumulate_l1(AccumulateArgs A) {
    accumulate_body<true>((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}
#else
#define k_accumulate_l1 k_accumulate<true>
#endif
// Accumulate levels >= 2: block b owns partial slots [b*ACC_TILE, (b+1)*ACC_TILE).  Segmented inclusive scan by key
// (keys are sorted, runs are contiguous), after which the last slot of every run holds the run's sum inside the tile.
// Output contract = accumulate_body<false> with L = ACC_TILE run by ONE thread over the same tile (that body is the host
// reference in tests/emul): closed runs go to their bucket, the at most two runs that cross the tile edge become the
// two partial slots of this block.
__global__ void __launch_bounds__(ACC_TILE) k_segscan(AccumulateArgs A) {
    __shared__ ShPoints sp;
    __shared__ uint32_t sk[ACC_TILE];
    __shared__ uint32_t okey[2];
    const int t = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * ACC_TILE;
    const size_t i = base + t;
    uint32_t raw = i < A.m ? A.keys[i] : KEY_SENTINEL;
    const uint32_t k = raw & KEY_MASK;
    xyzz p = (k != KEY_MASK && !(raw & KEY_FILL)) ? load_xyzz(&A.pts_in[i]) : xyzz_identity();
    sk[t] = k;
    sh_store(sp, t, p);
    if (t < 2) okey[t] = KEY_SENTINEL;
    __syncthreads();
    for (int d = 1; d < ACC_TILE; d <<= 1) {
        const bool take = t >= d && k != KEY_MASK && sk[t - d] == k;
        xyzz q;
        if (take) q = sh_load(sp, t - d);
        __syncthreads();
        if (take) {
            p = xyzz_add(q, p);
            sh_store(sp, t, p);
        }
        __syncthreads();
    }
    const bool run_end = t == ACC_TILE - 1 || sk[t + 1] != k;
    if (k != KEY_MASK && run_end) {
        const bool from_start = sk[0] == k;
        const bool to_end = t == ACC_TILE - 1;
        const bool left_open = from_start && base > 0 && (A.keys[base - 1] & KEY_MASK) == k;
        const bool right_open = to_end && base + ACC_TILE < A.m && (A.keys[base + ACC_TILE] & KEY_MASK) == k;
        if (!left_open && !right_open) {
            emit_bucket(A, k, p);
        } else if (from_start) {
            okey[0] = k;
            store_xyzz(&A.ppts[2 * (size_t)blockIdx.x], p);
            if (to_end) okey[1] = k | KEY_FILL;  // the whole tile is one run: the filler keeps it contiguous upstream
        } else {
            okey[1] = k;
            store_xyzz(&A.ppts[2 * (size_t)blockIdx.x + 1], p);
        }
    }
    __syncthreads();
    if (t < 2) A.pkeys[2 * (size_t)blockIdx.x + t] = okey[t];
}

// The stages behind the accumulate levels (bucket merge, bucket reduce, finish) live in depth_kernels.cu: they are bound
// by DEPTH (a chain of a few dozen group additions on a handful of warps) and are compiled with the field operations as
// calls instead of inline expansions (see there).
__global__ void __launch_bounds__(128) k_build_table(TableArgs A) {
    table_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}
__global__ void __launch_bounds__(128) k_table_chain(TableSlabArgs A) { table_chain_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
__global__ void __launch_bounds__(128) k_table_norm(TableSlabArgs A) { table_norm_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }

static inline unsigned grid_for(size_t threads, unsigned block) { return (unsigned)((threads + block - 1) / block); }

int launch_decompose(const DecomposeArgs& A, cudaStream_t st) {
    k_decompose<<<grid_for((size_t)A.g * A.n, 256), 256, 0, st>>>(A);
    COZK_CUDA(cudaGetLastError());
    return COZK_OK;
}

// Partial-slot buffers of the accumulate levels: two sets (`parity`), so that the levels >= 2 of one chunk of a streamed
// call can run beside level 1 of the next; inside a set, level l writes buffer l & 1 and reads the other.
static int ensure_accumulate_buffers(Device& D, const MsmPlan& P, int parity) {
    int rc;
    for (size_t lvl = 0; lvl < P.acc_entries.size(); ++lvl) {
        const int tile = P.acc_tile[lvl];
        const size_t T = (P.acc_entries[lvl] + tile - 1) / tile;
        if ((rc = D.pk[2 * parity + (lvl & 1)].ensure(2 * T * 4))) return rc;
        if ((rc = D.pp[2 * parity + (lvl & 1)].ensure(2 * T * sizeof(xyzz)))) return rc;
    }
    return COZK_OK;
}

// Accumulate stage: sorted (key, val) pairs -> bucket sums in `bucket_dst`, levels [lvl_begin, lvl_end) on stream st
// (lvl_end = 0: all).  Level 0 zeroes the bucket set first.
static int run_accumulate(Device& D, const MsmPlan& P, const uint32_t* keys, const uint32_t* vals, const affine* d_bases,
                          xyzz* bucket_dst, double* launches, cudaStream_t st = nullptr, int parity = 0, size_t lvl_begin = 0,
                          size_t lvl_end = 0) {
    if (!st) st = D.stream;
    if (!lvl_end) lvl_end = P.acc_entries.size();
    int rc;
    if ((rc = ensure_accumulate_buffers(D, P, parity))) return rc;
    if (lvl_begin == 0) COZK_CUDA(cudaMemsetAsync(bucket_dst, 0, P.total_buckets * sizeof(xyzz), st));
    for (size_t lvl = lvl_begin; lvl < lvl_end; ++lvl) {
        size_t m = P.acc_entries[lvl];
        const int tile = P.acc_tile[lvl];
        size_t T = (m + tile - 1) / tile;  // threads (serial body) or blocks (segmented scan)
        DevBuf& pk_out = D.pk[2 * parity + (lvl & 1)];
        DevBuf& pp_out = D.pp[2 * parity + (lvl & 1)];
        AccumulateArgs A{m,
                         lvl == 0 ? keys : D.pk[2 * parity + ((lvl - 1) & 1)].as<uint32_t>(),
                         vals,
                         d_bases,
                         lvl == 0 ? nullptr : D.pp[2 * parity + ((lvl - 1) & 1)].as<xyzz>(),
                         bucket_dst,
                         pk_out.as<uint32_t>(),
                         pp_out.as<xyzz>(),
                         (uint32_t)tile};
        if (lvl == 0) k_accumulate_l1<<<grid_for(T, COZK_ACC_BLOCK), COZK_ACC_BLOCK, 0, st>>>(A);
        else if (tile != ACC_TILE) k_accumulate<false><<<grid_for(T, 128), 128, 0, st>>>(A);
        else k_segscan<<<(unsigned)T, ACC_TILE, 0, st>>>(A);
        *launches += 1;
        COZK_CUDA(cudaGetLastError());
    }
    return COZK_OK;
}

// Batched-affine pre-reduction in front of the accumulate levels (affine.cu): how many halving rounds pay for this pair
// list.  A round needs runs to halve: after R rounds the average bucket still has to hold 4 entries.
static int affine_rounds_for(const Device& D, size_t m, size_t total_buckets) {
    int R = (int)D.affine_rounds;
    if (R <= 0 || m < (size_t)D.affine_min_pairs) return 0;
    if (R > AFF_MAX_ROUNDS) R = AFF_MAX_ROUNDS;
    while (R > 0 && (m >> R) < 4 * total_buckets) --R;
    return R;
}

// Level 1 of the accumulate stage on stream st, with the pre-reduction when it pays: sorted pairs -> partial slots of level 1
// (+ closed runs in bucket_dst).  *Pacc is the plan the levels >= 2 continue with; *AS the stage whose overflow lists have to
// join bucket_dst behind the last level (affine_overflow_adds; AS->rounds == 0: nothing to add).
static int run_level1(Device& D, const MsmPlan& P, const uint32_t* keys, const uint32_t* vals, const affine* d_bases, xyzz* bucket_dst,
                      double* launches, cudaStream_t st, int parity, MsmPlan* Pacc, AffineStage* AS) {
    *Pacc = P;
    AS->rounds = 0;
    const int R = affine_rounds_for(D, P.m, P.total_buckets);
    int rc;
    if (R > 0) {
        if ((rc = affine_reduce(D, st, parity, P.m, P.total_buckets, keys, vals, d_bases, R, AS, launches))) return rc;
        plan_set_pairs(*Pacc, AS->m);
        keys = AS->keys;
        vals = AS->vals;
        d_bases = AS->pts;
    }
    return run_accumulate(D, *Pacc, keys, vals, d_bases, bucket_dst, launches, st, parity, 0, 1);
}

// ------------------------------------------------------------------------------------------------ one group on one device
// Scalars are already on the device (contiguous staging or caller-owned vectors).  Results: g wire points in D.out.
static int run_group(Device& D, const MsmPlan& P, const affine* d_bases, const uint8_t* d_inf, const uint8_t* d_scalars,
                     const uint8_t* const* d_vec_ptrs, size_t vector_stride, size_t stride, int form, size_t table_stride,
                     size_t val_offset, double* launches, int phases = 3, bool merge = false, const DecomposeArgs* dom = nullptr,
                     const DecomposeArgs* rag = nullptr) {
    // rag: ragged group (only its rag_start / rag_base / total fields are read), or null
    // dom: dominant-digit layout of this group (only its dom_* / seg_* / totals_index fields are read), or null
    // phases: bit 0 = decompose + sort + accumulate into the buckets, bit 1 = bucket reduce + finish.  A streamed MSM
    // (one vector fed in point chunks while the next chunk is still on the PCIe bus) runs bit 0 once per chunk, with
    // merge = true from the second chunk on, and bit 1 once at the end.
    cudaStream_t st = D.stream;
    int rc;
    if (phases & 1) {
    if ((rc = D.keys_a.ensure(P.m * 4))) return rc;
    if ((rc = D.vals_a.ensure(P.m * 4))) return rc;
    if ((rc = D.keys_b.ensure(P.m * 4))) return rc;
    if ((rc = D.vals_b.ensure(P.m * 4))) return rc;
    if ((rc = D.buckets.ensure(P.total_buckets * sizeof(xyzz)))) return rc;
    if ((rc = D.out.ensure((size_t)P.g * sizeof(xyzz) + 256))) return rc;

    COZK_CUDA(cudaEventRecord(D.ev[1], st));
    // 1 + 2: decompose and sort.  Plain layout: the pairs are produced inside the first pass of the sort (born
    // partitioned by their low key digit, never stored unsorted).  Dominant-digit layout: compacted segments from the
    // decompose kernel, then generic passes.  Stage times: "decompose" = up to and including the first sort pass.
    DecomposeArgs DA{d_scalars, d_vec_ptrs, vector_stride, stride, form, P.n, P.g, P.c, P.W, d_inf,
                     D.keys_a.as<uint32_t>(), D.vals_a.as<uint32_t>(), P.Wb, table_stride, val_offset};
    uint32_t *sorted_keys = nullptr, *sorted_vals = nullptr;
    if (rag) {
        DA.rag_start = rag->rag_start;
        DA.rag_base = rag->rag_base;
        DA.total = rag->total;
    }
    if (dom) {
        DA.dom_mode = dom->dom_mode;
        DA.dom_cand = dom->dom_cand;
        DA.seg_off = dom->seg_off;
        DA.seg_cursor = dom->seg_cursor;
        DA.seg_len = dom->seg_len;
        DA.totals_index = dom->totals_index;
        k_decompose<<<grid_for((size_t)P.g * P.n, 256), 256, 0, st>>>(DA);
        *launches += 1;
        COZK_CUDA(cudaGetLastError());
        COZK_CUDA(cudaEventRecord(D.ev[2], st));
        if ((rc = sort_pairs(D, st, nullptr, P.m, P.sort_bits, &sorted_keys, &sorted_vals, launches, nullptr))) return rc;
    } else {
        if ((rc = sort_pairs(D, st, &DA, P.m, P.sort_bits, &sorted_keys, &sorted_vals, launches, D.ev[2]))) return rc;
    }
    COZK_CUDA(cudaEventRecord(D.ev[3], st));

    // 3 accumulate, level by level
    xyzz* bucket_dst = D.buckets.as<xyzz>();
    if (merge) {
        if ((rc = D.buckets2[0].ensure(P.total_buckets * sizeof(xyzz)))) return rc;
        bucket_dst = D.buckets2[0].as<xyzz>();
    }
    {
        MsmPlan Pacc;
        AffineStage AS;
        if ((rc = run_level1(D, P, sorted_keys, sorted_vals, d_bases, bucket_dst, launches, st, 0, &Pacc, &AS))) return rc;
        if (Pacc.acc_entries.size() > 1 && (rc = run_accumulate(D, Pacc, nullptr, nullptr, nullptr, bucket_dst, launches, st, 0, 1, 0))) return rc;
        if (AS.rounds && (rc = affine_overflow_adds(st, AS, bucket_dst, launches))) return rc;
    }
    if (merge) {
        MergeArgs MA{D.buckets.as<xyzz>(), bucket_dst, P.total_buckets};
        launch_merge(MA, grid_for(P.total_buckets, 128), st);
        *launches += 1;
        COZK_CUDA(cudaGetLastError());
    }
    COZK_CUDA(cudaEventRecord(D.ev[4], st));
    }  // phase bit 0
    if (!(phases & 2)) return COZK_OK;

    // 4 bucket reduce.  Row / column form: two block-cooperative tree sums.  Group form: group running sums, then NS plain
    // sums per window.
    size_t windows = (size_t)P.g * P.Wb;
    size_t nsums = windows * P.NS;
    xyzz* cur = nullptr;
    if (P.reduce_2d) {
        const size_t rc_n = windows * (((size_t)1 << P.hi_bits) + ((size_t)1 << P.lo_bits));
        if (rc_n > 0x7FFFFFFFull || nsums > 0x7FFFFFFFull) {
            set_error("internal: too many bucket sets for the row / column reduce");
            return COZK_ERR_INVALID_ARG;
        }
        if ((rc = D.rs[0].ensure(rc_n * sizeof(xyzz)))) return rc;
        if ((rc = D.rs[1].ensure(nsums * sizeof(xyzz)))) return rc;
        RowColArgs RA{D.buckets.as<xyzz>(), D.rs[0].as<xyzz>(), P.lo_bits, P.hi_bits, rc_n};
        launch_rowcol(RA, st);
        COZK_CUDA(cudaGetLastError());
        MaskSumArgs MA{D.rs[0].as<xyzz>(), D.rs[1].as<xyzz>(), P.lo_bits, P.hi_bits, P.NS, nsums};
        launch_masksum(MA, st);
        COZK_CUDA(cudaGetLastError());
        *launches += 2;
        cur = D.rs[1].as<xyzz>();
    } else {
    size_t groups = windows * P.G;
    if ((rc = D.rs[0].ensure(groups * sizeof(xyzz)))) return rc;
    if ((rc = D.rw[0].ensure(groups * sizeof(xyzz)))) return rc;
    GroupArgs GA{D.buckets.as<xyzz>(), D.rs[0].as<xyzz>(), D.rw[0].as<xyzz>(), P.group_l, groups};
    launch_group(GA, grid_for(groups, 64), st);
    *launches += 1;
    COZK_CUDA(cudaGetLastError());
    if ((rc = D.rs[1].ensure(nsums * P.sum_chunks * sizeof(xyzz)))) return rc;
    if ((rc = D.rw[1].ensure(nsums * sizeof(xyzz)))) return rc;
    TreeSumArgs TA{D.rs[0].as<xyzz>(), D.rw[0].as<xyzz>(), D.rs[1].as<xyzz>(), P.G, P.NS, P.sum_chunk, P.sum_chunks, 1};
    launch_treesum(TA, (unsigned)(windows * P.NS * P.sum_chunks), st);
    *launches += 1;
    COZK_CUDA(cudaGetLastError());
    cur = D.rs[1].as<xyzz>();
    if (P.sum_chunks > 1) {
        TreeSumArgs TB{cur, nullptr, D.rw[1].as<xyzz>(), P.sum_chunks, P.NS, P.sum_chunks, 1, 0};
        launch_treesum(TB, (unsigned)(windows * P.NS), st);
        *launches += 1;
        COZK_CUDA(cudaGetLastError());
        cur = D.rw[1].as<xyzz>();
    }
    }
    COZK_CUDA(cudaEventRecord(D.ev[5], st));

    // 5 finish: bit-position Horner + inversion.  Few vectors: on the host (a CPU core runs this serial chain an order
    // of magnitude faster than one GPU thread); large batches: one GPU thread per vector, all in parallel.
    if (P.g <= HOST_FINISH_MAX) {
        D.host_sums.resize(nsums);
        COZK_CUDA(cudaMemcpyAsync(D.host_sums.data(), cur, nsums * sizeof(xyzz), cudaMemcpyDeviceToHost, st));
        D.finish_on_host = true;
    } else {
        // the Horner pass on the device, one thread per vector; the sums stay un-normalised (D.out: g XYZZ points) and the
        // caller normalises them on the host (fetch_device_finish / host_normalize_batch: one inversion for the whole batch)
        FinishArgs F{cur, P.g, P.Wb, P.c, P.NS, P.log_l, nullptr, D.out.as<xyzz>()};
        launch_finish(F, grid_for(P.g, 32), st);
        *launches += 1;
        COZK_CUDA(cudaGetLastError());
        D.finish_on_host = false;
    }
    return COZK_OK;
}

static void add_stage_times(Device& D, int first = 1, int last = 5) {
    float ms;
    for (int s = first; s <= last; ++s) {
        if (cudaEventElapsedTime(&ms, D.ev[s], D.ev[s + 1]) == cudaSuccess) D.stats[s] += ms;
    }
}

// host-side sum of wire points (used to combine point-range shards): out = sum of count 72-byte points
static void host_sum(const uint8_t* pts, size_t count, uint8_t* out) {
    xyzz acc = xyzz_identity();
    for (size_t i = 0; i < count; ++i) acc = xyzz_add(acc, xyzz_from_wire(pts + 72 * i));
    xyzz_to_wire(acc, out);
}

constexpr size_t MAX_POINTS_PER_PASS = (size_t)1 << 26;  // default of option "max_points_per_pass"

// Dominant-digit analysis of one group (scalars already on the device): candidate digits from every vector's first
// scalar, a look at the first 1024 scalars of every vector, and - only if that sample shows a dominated window - the full
// count.  Fills `dom` and shrinks the plan's pair count when at least one segment is special; returns false otherwise.
// g un-normalised sums -> 72-byte wire points with ONE inversion (Montgomery's trick over the ZZ * ZZZ of the finite ones);
// the same canonical bytes as xyzz_to_wire point by point
static void host_normalize_batch(const xyzz* pts, size_t g, uint8_t* out72) {
    std::vector<fq> den(g), pre(g);
    fq run = fq_one();
    for (size_t v = 0; v < g; ++v) {
        pre[v] = run;
        if (xyzz_is_identity(pts[v])) continue;
        den[v] = fq_mul(pts[v].ZZ, pts[v].ZZZ);
        run = fq_mul(run, den[v]);
    }
    fq inv = fq_inv(run);
    for (size_t v = g; v-- > 0;) {
        uint8_t* o = out72 + 72 * v;
        if (xyzz_is_identity(pts[v])) {
            memset(o, 0, 72);
            o[64] = 1;
            continue;
        }
        const fq I = fq_mul(inv, pre[v]);  // 1 / (ZZ * ZZZ) of point v
        inv = fq_mul(inv, den[v]);
        const fq x = fq_mul(pts[v].X, fq_mul(I, pts[v].ZZZ)), y = fq_mul(pts[v].Y, fq_mul(I, pts[v].ZZ));
        memcpy(o, x.v, 32);
        memcpy(o + 32, y.v, 32);
        memset(o + 64, 0, 8);
    }
}
// After run_group chose the on-device finish: bring the g un-normalised sums back (enqueued; valid after the stream's next
// synchronisation) ...
static int fetch_device_finish(Device& D, size_t g, cudaStream_t st) {
    D.host_sums.resize(g);
    COZK_CUDA(cudaMemcpyAsync(D.host_sums.data(), D.out.p, g * sizeof(xyzz), cudaMemcpyDeviceToHost, st));
    return COZK_OK;
}

static int analyse_dominant(Device& D, MsmPlan& P, const uint8_t* d_scalars, const uint8_t* const* d_vec_ptrs, size_t vector_stride,
                            size_t stride, int form, size_t totals_index, DecomposeArgs* dom, bool* use, double* launches) {
    *use = false;
    const size_t segs = (size_t)P.g * P.W;
    int rc;
    if ((rc = D.dom_cand.ensure(segs * 4))) return rc;
    if ((rc = D.dom_counts.ensure(2 * segs * 4))) return rc;
    DomArgs A{{d_scalars, d_vec_ptrs, vector_stride, stride, form, P.n, P.g, P.c, P.W, nullptr, nullptr, nullptr, P.Wb, 0, 0},
              D.dom_cand.as<int32_t>(), D.dom_counts.as<uint32_t>(), D.dom_counts.as<uint32_t>() + segs, 0};
    cudaStream_t st = D.stream;
    k_dom_cand<<<grid_for(P.g, 32), 32, 0, st>>>(A);
    COZK_CUDA(cudaGetLastError());
    D.h_dom_cand.resize(segs);
    D.h_dom_counts.resize(2 * segs);
    const size_t sample = std::min<size_t>(P.n, 1024);
    for (int round = 0; round < 2; ++round) {
        A.count_n = round == 0 ? sample : P.n;
        COZK_CUDA(cudaMemsetAsync(D.dom_counts.p, 0, 2 * segs * 4, st));
        {
            // enough blocks to fill the device, never more than the scalars need
            unsigned per_vec = std::min<unsigned>(grid_for(A.count_n, 256), std::max<unsigned>(1u, (unsigned)(D.sm_count * 8 / P.g) + 1));
            k_dom_count<<<dim3(per_vec, P.g), 256, 0, st>>>(A);
        }
        *launches += 1;
        COZK_CUDA(cudaGetLastError());
        COZK_CUDA(cudaMemcpyAsync(D.h_dom_counts.data(), D.dom_counts.p, 2 * segs * 4, cudaMemcpyDeviceToHost, st));
        if (round == 0) COZK_CUDA(cudaMemcpyAsync(D.h_dom_cand.data(), D.dom_cand.p, segs * 4, cudaMemcpyDeviceToHost, st));
        COZK_CUDA(cudaStreamSynchronize(st));
        size_t m = dom_layout(D.h_dom_cand.data(), D.h_dom_counts.data(), D.h_dom_counts.data() + segs, segs, A.count_n, D.h_dom_mode,
                              D.h_dom_off, D.h_dom_len);
        if (m == 0) return COZK_OK;            // nothing dominated (in the sample, or in the whole vectors)
        if (A.count_n == P.n) {
            plan_set_pairs(P, m);
            break;
        }
    }
    if ((rc = D.dom_mode.ensure(segs * 4))) return rc;
    if ((rc = D.dom_off.ensure(segs * 8))) return rc;
    if ((rc = D.dom_len.ensure(segs * 8))) return rc;
    if ((rc = D.dom_cursor.ensure(segs * 4))) return rc;
    COZK_CUDA(cudaMemcpyAsync(D.dom_mode.p, D.h_dom_mode.data(), segs * 4, cudaMemcpyHostToDevice, st));
    COZK_CUDA(cudaMemcpyAsync(D.dom_off.p, D.h_dom_off.data(), segs * 8, cudaMemcpyHostToDevice, st));
    COZK_CUDA(cudaMemcpyAsync(D.dom_len.p, D.h_dom_len.data(), segs * 8, cudaMemcpyHostToDevice, st));
    COZK_CUDA(cudaMemsetAsync(D.dom_cursor.p, 0, segs * 4, st));
    dom->dom_mode = D.dom_mode.as<uint32_t>();
    dom->dom_cand = D.dom_cand.as<int32_t>();
    dom->seg_off = D.dom_off.as<uint64_t>();
    dom->seg_cursor = D.dom_cursor.as<uint32_t>();
    dom->seg_len = D.dom_len.as<uint64_t>();
    dom->totals_index = totals_index;
    *use = true;
    return COZK_OK;
}

// All k vectors over bases [offset, offset+n) on ONE device.  host_scalars xor dev_scalars.
static int run_on_device(cozk_ctx* ctx, int dev_index, const SrsEntry& S, size_t offset, size_t n,
                         const void* const* host_scalars, const void* const* dev_scalars, size_t k, size_t stride, int form,
                         unsigned max_bits, uint8_t* out, long stream_min = -1) {
    // stream_min: points from which a single host vector is streamed in chunks; -1 = option "stream_min_points"
    if (stream_min < 0) stream_min = ctx->opt_stream_min_points;
    Device& D = *ctx->devs[dev_index];
    if (!S.bases(dev_index)) {
        set_error("this SRS does not live on the device the call was sent to");
        return COZK_ERR_INVALID_ARG;
    }
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    D.sort_digit_bits = ctx->opt_sort_digit_bits;
    D.affine_rounds = ctx->opt_affine_rounds;
    D.affine_min_pairs = ctx->opt_affine_min_pairs;
    for (double& s : D.stats) s = 0;
    double launches = 0;
    uint32_t bits = (max_bits == 0 || max_bits > 254) ? 254 : max_bits;
    const size_t max_buckets = (size_t)1 << 25;  // 4 GiB of XYZZ buckets per group
    // a call longer than this runs as several passes over point ranges whose results are added on the host
    const size_t per_pass = (size_t)std::max<long>(1, ctx->opt_max_points_per_pass);
    size_t passes = (n + per_pass - 1) / per_pass;
    std::vector<uint8_t> partial(passes > 1 ? passes * k * 72 : 0);
    COZK_CUDA(cudaEventRecord(D.ev[0], D.stream));
    double plan_mults = 0, plan_pairs = 0, host_finish_ms = 0;
    uint32_t last_c = 0, last_W = 0;
    // k_accumulate<true>: 4 blocks of 128 threads per SM are resident
    const AccTuning acc_tuning{(size_t)D.sm_count * 512, (int)ctx->opt_acc_chunk, (int)ctx->opt_acc_chunk_up, (int)ctx->opt_group_l,
                               (int)ctx->opt_reduce_2d};

    for (size_t pass = 0; pass < passes; ++pass) {
        size_t lo = pass * per_pass;
        size_t pn = std::min(per_pass, n - lo);
        const uint8_t* d_inf = S.inf(dev_index) ? S.inf(dev_index) + offset + lo : nullptr;
        uint8_t* pass_out = passes > 1 ? partial.data() + pass * k * 72 : out;

        // With a precomputed table (2^(c*w) * P rows built at registration) all windows share one bucket set; use it
        // when the cost model says so (a short prefix of a long SRS is cheaper with its own, smaller window).
        uint32_t table_c = 0;
        if (S.table_c && !ctx->opt_window) {
            MsmPlan with_table = make_plan(pn, 1, bits, max_buckets, 0, S.table_c);
            MsmPlan without = make_plan(pn, 1, bits, max_buckets, 0, 0);
            if (with_table.W <= S.table_W && with_table.model_cost() <= without.model_cost()) table_c = S.table_c;
        }
        const affine* d_bases = table_c ? S.bases(dev_index) : S.bases(dev_index) + offset + lo;
        const size_t table_stride = table_c ? S.n : 0, val_offset = table_c ? offset + lo : 0;

        // One long vector (host memory or device-resident): feed it in point chunks through ONE bucket set.  From host memory
        // the H2D copy of chunk i + 1 overlaps the work on chunk i; in both cases the decompose + sort of chunk i + 1 and the
        // latency-bound upper accumulate levels of chunk i run BESIDE the level-1 kernels, which follow each other without a
        // gap (pipeline below); reduce and finish run once.
        // chunks: option "stream_chunks", or (0 = auto) 2.  Measured end to end from pinned host memory, 2 / 3 / 4 / 6 chunks:
        // 2^20 3.61 / 3.67 / 3.75 / 4.06 ms (round-2 three-stream form: 3.70 / 3.84 / 3.98), 2^22 11.62 / 11.65 / 11.89 / 12.33
        // (12.15), 2^24 40.6 / 40.7 / 40.8 / 41.9 (43.2): every chunk more costs another set of upper levels and a merge.
        const long opt_chunks = ctx->opt_stream_chunks;
        const long stream_chunks = opt_chunks ? opt_chunks : 2;
        const long chunk_min = host_scalars ? stream_min : (long)ctx->opt_chunk_min_points;
        bool stream_it = k == 1 && stream_chunks > 1 && chunk_min > 0 && pn >= (size_t)chunk_min;
        const uint8_t* resident0 = (!host_scalars && stream_it) ? reinterpret_cast<const uint8_t*>(dev_scalars[0]) + lo * stride : nullptr;
        if (stream_it) {
            // The dominant-digit mode (constant and nearly constant share vectors: a 2^22 constant vector in 0.8 ms instead of
            // 10) needs the whole vector on the device before it can lay the pairs out, so it does not run in chunks.  Where its
            // preconditions hold, look at the head of the vector first (1024 scalars: a 32 KB copy and two tiny kernels):
            // if a window is dominated there, this call takes the one-shot path below.
            uint32_t tl = 0;
            while (tl < S.total_levels && (S.n >> tl) != pn) ++tl;
            const bool totals_addressable = (double)S.table_W * (double)S.n + (double)(S.total_levels + 1) * S.table_W < 2147483647.0;
            if (ctx->opt_dominant && tl < S.total_levels && ((S.total_ok >> tl) & 1) && (S.n >> tl << tl) == S.n && !d_inf &&
                totals_addressable && passes == 1 && offset == 0 && (double)pn >= (double)ctx->opt_dominant_min_points) {
                const size_t head = std::min<size_t>(pn, 1024);
                int rc;
                const uint8_t* head_src = resident0;
                if (host_scalars) {
                    if ((rc = D.scalars[0].ensure((head - 1) * stride + 32 + 256))) return rc;
                    COZK_CUDA(cudaMemcpyAsync(D.scalars[0].p, reinterpret_cast<const uint8_t*>(host_scalars[0]) + lo * stride,
                                              (head - 1) * stride + 32, cudaMemcpyHostToDevice, D.stream));
                    head_src = D.scalars[0].as<uint8_t>();
                }
                MsmPlan Ph = make_plan(head, 1, bits, max_buckets, (uint32_t)ctx->opt_window, table_c, acc_tuning);
                if (!table_c) Ph = make_plan(head, 1, bits, max_buckets, make_plan(pn, 1, bits, max_buckets, (uint32_t)ctx->opt_window, 0).c, 0, acc_tuning);
                DecomposeArgs peek = {};
                bool dominated = false;
                rc = analyse_dominant(D, Ph, head_src, nullptr, 0, stride, form, 0, &peek, &dominated, &launches);
                if (rc) return rc;
                if (dominated) stream_it = false;
            }
        }
        if (stream_it) {
            // (as a lambda: a failure in the middle must not leave work of this call pending on the side streams)
            const int crc = [&]() -> int {
            // Pipeline without host synchronisation:
            //   copy   H2D of chunk i + 1 into staging slot (i + 1) & 1 (waits until the sort of chunk i - 1 has consumed it)
            //   prep   decompose + sort of chunk i into sort-buffer set i & 1 (high priority: its blocks get the next SM that
            //          frees up, so the sorted pairs of chunk i + 1 are ready long before level 1 of chunk i ends)
            //   l1[i & 1]  LEVEL 1 of the accumulate stage of chunk i - the throughput-bound part.  Alternating streams: the blocks
            //          of chunk i + 1 fill the SMs that chunk i's last blocks leave, no tail, no gap
            //   aux    levels >= 2 and the bucket merge of chunk i (chains of a few dozen additions on a handful of warps:
            //          bound by latency) - beside level 1 of chunk i + 1 instead of in front of it.
            // Chunk 0 accumulates into the bucket set itself, chunk i > 0 into scratch set i & 1, merged on aux in order.
            // Chunk lengths: the FIRST chunk is half of an equal share - nothing runs beside its copy and its sort - the others
            // share the rest.
            const size_t C = (size_t)stream_chunks;
            std::vector<size_t> cstart;  // chunk ci covers [cstart[ci], cstart[ci + 1])
            {
                // option "stream_first_pct": the first chunk's share of an equal share, in percent (default 50)
                const size_t pct = (size_t)std::max<long>(1, std::min<long>(100 * (long)C, ctx->opt_stream_first_pct));
                const size_t first = std::min(pn, std::max<size_t>(32, ((pn * pct / (100 * C)) + 31) & ~(size_t)31));
                const size_t rest = C > 1 ? ((((pn - first) + (C - 1) - 1) / (C - 1)) + 31) & ~(size_t)31 : 0;
                cstart.push_back(0);
                for (size_t at = first; at < pn; at += std::max<size_t>(rest, 32)) cstart.push_back(at);
                cstart.push_back(pn);
            }
            const size_t chunks = cstart.size() - 1;
            size_t cn_max = 0;
            for (size_t ci = 0; ci < chunks; ++ci) cn_max = std::max(cn_max, cstart[ci + 1] - cstart[ci]);
            const uint32_t cfix = table_c ? 0 : make_plan(pn, 1, bits, max_buckets, (uint32_t)ctx->opt_window, 0).c;
            const uint32_t cuse = cfix ? cfix : (uint32_t)ctx->opt_window;
            const uint8_t* src0 = host_scalars ? reinterpret_cast<const uint8_t*>(host_scalars[0]) + lo * stride : nullptr;
            int rc;
            // every buffer at its final size before the first launch: cudaFree / cudaMalloc would synchronise the device
            const MsmPlan Pmax = make_plan(cn_max, 1, bits, max_buckets, cuse, table_c, acc_tuning);
            if (host_scalars)
                for (int sl = 0; sl < 2; ++sl)
                    if ((rc = D.scalars[sl].ensure((cn_max - 1) * stride + 32 + 256))) return rc;
            for (DevBuf* b : {&D.keys_a, &D.vals_a, &D.keys_b, &D.vals_b, &D.keys_a2, &D.vals_a2, &D.keys_b2, &D.vals_b2})
                if ((rc = b->ensure(Pmax.m * 4))) return rc;
            if ((rc = D.buckets.ensure(Pmax.total_buckets * sizeof(xyzz))) || (rc = D.out.ensure(72 + 256))) return rc;
            for (int par = 0; par < 2; ++par) {
                if (chunks > 1 && (rc = D.buckets2[par].ensure(Pmax.total_buckets * sizeof(xyzz)))) return rc;
                if ((rc = ensure_accumulate_buffers(D, Pmax, par))) return rc;
            }
            if (!D.aux_stream) COZK_CUDA(cudaStreamCreateWithFlags(&D.aux_stream, cudaStreamNonBlocking));
            if (!D.l1b_stream) COZK_CUDA(cudaStreamCreateWithFlags(&D.l1b_stream, cudaStreamNonBlocking));
            if (!D.prep_stream) {
                int prio_low = 0, prio_high = 0;
                COZK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
                COZK_CUDA(cudaStreamCreateWithPriority(&D.prep_stream, cudaStreamNonBlocking, prio_high));
            }
            while (D.chunk_ev.size() < 6 * chunks) {
                cudaEvent_t e;
                COZK_CUDA(cudaEventCreate(&e));
                D.chunk_ev.push_back(e);
            }
            // events of chunk ci: 0 sort starts, 1 first sort pass done, 2 sorted (staging slot free), 3 level 1 done (sort-buffer
            // set free), 4 chunk done (partial-slot set and scratch bucket set free), 5 level 1 starts
            auto EV = [&](size_t ci, int which) { return D.chunk_ev[6 * ci + which]; };
            auto stage_chunk = [&](size_t ci) -> int {
                const int slot = (int)(ci & 1);
                const size_t clo = cstart[ci], cn = cstart[ci + 1] - clo;
                if (ci >= 2) COZK_CUDA(cudaStreamWaitEvent(D.copy_stream, EV(ci - 2, 2), 0));
                COZK_CUDA(cudaMemcpyAsync(D.scalars[slot].p, src0 + clo * stride, (cn - 1) * stride + 32, cudaMemcpyHostToDevice,
                                          D.copy_stream));
                COZK_CUDA(cudaEventRecord(D.copy_done[slot], D.copy_stream));
                return COZK_OK;
            };
            // everything below is ordered behind what the caller's stream has done so far (the peek above, earlier calls)
            COZK_CUDA(cudaEventRecord(D.ev[1], D.stream));
            for (cudaStream_t s : {D.prep_stream, D.l1b_stream, D.aux_stream, D.copy_stream}) COZK_CUDA(cudaStreamWaitEvent(s, D.ev[1], 0));
            if (host_scalars && (rc = stage_chunk(0))) return rc;
            MsmPlan P;
            for (size_t ci = 0; ci < chunks; ++ci) {
                const int slot = (int)(ci & 1), par = (int)(ci & 1);
                const size_t clo = cstart[ci], cn = cstart[ci + 1] - clo;
                cudaStream_t l1 = par ? D.l1b_stream : D.stream;
                P = make_plan(cn, 1, bits, max_buckets, cuse, table_c, acc_tuning);
                plan_mults += P.field_mults();
                plan_pairs += (double)P.m;
                last_c = P.c;
                last_W = P.W;
                if (host_scalars) COZK_CUDA(cudaStreamWaitEvent(D.prep_stream, D.copy_done[slot], 0));
                if (ci >= 2) COZK_CUDA(cudaStreamWaitEvent(D.prep_stream, EV(ci - 2, 3), 0));  // level 1 of chunk ci - 2 read this set
                COZK_CUDA(cudaEventRecord(EV(ci, 0), D.prep_stream));
                if (host_scalars && ci + 1 < chunks && (rc = stage_chunk(ci + 1))) return rc;
                const affine* cb = table_c ? d_bases : d_bases + clo;
                const uint8_t* csrc = host_scalars ? D.scalars[slot].as<uint8_t>() : resident0 + clo * stride;
                DecomposeArgs DA{csrc, nullptr, 0, stride, form, P.n, P.g, P.c, P.W, d_inf ? d_inf + clo : nullptr,
                                 nullptr, nullptr, P.Wb, table_stride, val_offset + (table_c ? clo : 0)};
                uint32_t *sk = nullptr, *sv = nullptr;
                if ((rc = sort_pairs(D, D.prep_stream, &DA, P.m, P.sort_bits, &sk, &sv, &launches, EV(ci, 1), par))) return rc;
                COZK_CUDA(cudaEventRecord(EV(ci, 2), D.prep_stream));
                xyzz* dst = ci == 0 ? D.buckets.as<xyzz>() : D.buckets2[par].as<xyzz>();
                COZK_CUDA(cudaStreamWaitEvent(l1, EV(ci, 2), 0));
                // the partial-slot set and the scratch bucket set of this parity were last used by chunk ci - 2
                if (ci >= 2) COZK_CUDA(cudaStreamWaitEvent(l1, EV(ci - 2, 4), 0));
                COZK_CUDA(cudaEventRecord(EV(ci, 5), l1));
                MsmPlan Pacc;
                AffineStage AS;
                if ((rc = run_level1(D, P, sk, sv, cb, dst, &launches, l1, par, &Pacc, &AS))) return rc;
                COZK_CUDA(cudaEventRecord(EV(ci, 3), l1));
                COZK_CUDA(cudaStreamWaitEvent(D.aux_stream, EV(ci, 3), 0));
                if (Pacc.acc_entries.size() > 1 && (rc = run_accumulate(D, Pacc, nullptr, nullptr, cb, dst, &launches, D.aux_stream, par, 1, 0)))
                    return rc;
                if (AS.rounds && (rc = affine_overflow_adds(D.aux_stream, AS, dst, &launches))) return rc;
                if (ci > 0) {
                    MergeArgs MA{D.buckets.as<xyzz>(), dst, P.total_buckets};
                    launch_merge(MA, grid_for(P.total_buckets, 128), D.aux_stream);
                    launches += 1;
                    COZK_CUDA(cudaGetLastError());
                }
                COZK_CUDA(cudaEventRecord(EV(ci, 4), D.aux_stream));
            }
            COZK_CUDA(cudaStreamWaitEvent(D.stream, EV(chunks - 1, 4), 0));  // aux runs in order: the last chunk closes them all
            COZK_CUDA(cudaEventRecord(D.ev[4], D.stream));
            rc = run_group(D, P, d_bases, d_inf, nullptr, nullptr, 0, stride, form, table_stride, val_offset, &launches, 2, false);
            if (rc) return rc;
            if (!D.finish_on_host && (rc = fetch_device_finish(D, 1, D.stream))) return rc;
            COZK_CUDA(cudaEventRecord(D.ev[6], D.stream));
            COZK_CUDA(cudaStreamSynchronize(D.stream));
            add_stage_times(D, 4, 5);
            // Stage times.  The stages of different chunks overlap, so they are reported as what each adds to the critical
            // path: "decompose" = the first chunk up to its first sort pass, "sort" = the rest of the first chunk's sort,
            // "accumulate" = from the first level-1 launch to the end of the last chunk's upper levels and merge.
            {
                float ms;
                if (cudaEventElapsedTime(&ms, D.ev[1], EV(0, 1)) == cudaSuccess) D.stats[1] += ms;
                if (cudaEventElapsedTime(&ms, EV(0, 1), EV(0, 2)) == cudaSuccess) D.stats[2] += ms;
                if (cudaEventElapsedTime(&ms, EV(0, 2), EV(chunks - 1, 4)) == cudaSuccess) D.stats[3] += ms;
                if (getenv("COZK_CHUNK_TRACE")) {  // timeline of the chunks on stderr, ms since the call entered the pipeline
                    auto at = [&](cudaEvent_t e) {
                        float t = -1;
                        cudaEventElapsedTime(&t, D.ev[1], e);
                        return t;
                    };
                    for (size_t ci = 0; ci < chunks; ++ci)
                        fprintf(stderr, "[chunk %zu: %zu points] sort %.3f .. first pass %.3f .. %.3f | level 1 %.3f .. %.3f | upper levels + merge .. %.3f\n",
                                ci, cstart[ci + 1] - cstart[ci], at(EV(ci, 0)), at(EV(ci, 1)), at(EV(ci, 2)), at(EV(ci, 5)), at(EV(ci, 3)),
                                at(EV(ci, 4)));
                    fprintf(stderr, "[reduce] %.3f .. %.3f | finish copy .. %.3f\n", at(D.ev[4]), at(D.ev[5]), at(D.ev[6]));
                }
            }
            if (D.finish_on_host) {
                auto h0 = std::chrono::steady_clock::now();
                FinishArgs F{D.host_sums.data(), P.g, P.Wb, P.c, P.NS, P.log_l, pass_out};
                finish_body(0, F);
                host_finish_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
            } else {
                host_normalize_batch(D.host_sums.data(), 1, pass_out);
            }
            return COZK_OK;
            }();
            if (crc) {
                cudaDeviceSynchronize();
                return crc;
            }
            continue;
        }

        // vectors per group: bounded by the pair budget
        MsmPlan probe = make_plan(pn, 1, bits, max_buckets, (uint32_t)ctx->opt_window, table_c);
        size_t per_vec = probe.m;
        size_t gmax = std::max<size_t>(1, (size_t)ctx->opt_group_pairs / std::max<size_t>(per_vec, 1));
        gmax = std::min<size_t>(gmax, 4096);
        // ... and by the bucket budget: a fixed window (table or option) bounds g directly, a free one is searched
        if (table_c || ctx->opt_window) gmax = std::min<size_t>(gmax, std::max<size_t>(1, max_buckets / ((size_t)probe.Wb * probe.B)));
        else gmax = max_group_for(pn, (uint32_t)std::min<size_t>(gmax, k), bits, max_buckets);
        size_t vstride = ((pn - 1) * stride + 32 + 255) & ~(size_t)255;

        size_t ngroups = (k + gmax - 1) / gmax;
        // stage group 0's scalars, then overlap the copy of group i+1 with the compute of group i
        auto stage = [&](size_t gi) -> int {
            size_t v0 = gi * gmax, g = std::min(gmax, k - v0);
            int slot = (int)(gi & 1);
            if (host_scalars) {
                int rc = D.scalars[slot].ensure(g * vstride);
                if (rc) return rc;
                for (size_t v = 0; v < g; ++v) {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(host_scalars[v0 + v]) + lo * stride;
                    COZK_CUDA(cudaMemcpyAsync(D.scalars[slot].as<uint8_t>() + v * vstride, src, (pn - 1) * stride + 32,
                                              cudaMemcpyHostToDevice, D.copy_stream));
                }
            } else if (g == 1) {
                // one device-resident vector: its pointer travels in the kernel arguments, nothing to stage (no pointer-table copy,
                // no host synchronisation in front of the call)
            } else {
                int rc = D.vec_ptrs.ensure(2 * 4096 * sizeof(void*));
                if (rc) return rc;
                std::vector<const uint8_t*> ptrs(g);
                for (size_t v = 0; v < g; ++v) ptrs[v] = reinterpret_cast<const uint8_t*>(dev_scalars[v0 + v]) + lo * stride;
                COZK_CUDA(cudaMemcpyAsync(D.vec_ptrs.as<const uint8_t*>() + slot * 4096, ptrs.data(), g * sizeof(void*),
                                          cudaMemcpyHostToDevice, D.copy_stream));
                COZK_CUDA(cudaStreamSynchronize(D.copy_stream));  // ptrs goes out of scope
            }
            COZK_CUDA(cudaEventRecord(D.copy_done[slot], D.copy_stream));
            return COZK_OK;
        };
        int rc = stage(0);
        if (rc) return rc;
        for (size_t gi = 0; gi < ngroups; ++gi) {
            size_t v0 = gi * gmax, g = std::min(gmax, k - v0);
            int slot = (int)(gi & 1);
            MsmPlan P = make_plan(pn, (uint32_t)g, bits, max_buckets, (uint32_t)ctx->opt_window, table_c, acc_tuning);
            last_c = P.c;
            last_W = P.W;
            COZK_CUDA(cudaStreamWaitEvent(D.stream, D.copy_done[slot], 0));
            if (gi + 1 < ngroups) {
                // the other staging slot was last read by group gi-1, which has been synchronised below
                if ((rc = stage(gi + 1))) return rc;
            }
            const bool one_resident = !host_scalars && g == 1;
            const uint8_t* g_scalars = host_scalars ? D.scalars[slot].as<uint8_t>()
                                       : one_resident ? reinterpret_cast<const uint8_t*>(dev_scalars[v0]) + lo * stride : nullptr;
            const uint8_t* const* g_ptrs = (host_scalars || one_resident) ? nullptr : D.vec_ptrs.as<const uint8_t*>() + slot * 4096;
            DecomposeArgs dom = {};
            bool use_dom = false;
            // the whole SRS or a power-of-two prefix of it: the sums of those point ranges were computed at registration
            uint32_t tl = 0;
            while (tl < S.total_levels && (S.n >> tl) != pn) ++tl;
            // (a total's index shares the 31 index bits of `val` with the table entries)
            const bool totals_addressable = (double)S.table_W * (double)S.n + (double)(S.total_levels + 1) * S.table_W < 2147483647.0;
            if (ctx->opt_dominant && tl < S.total_levels && ((S.total_ok >> tl) & 1) && (S.n >> tl << tl) == S.n && !d_inf &&
                totals_addressable && passes == 1 && offset == 0 &&
                (double)g * (double)pn >= (double)ctx->opt_dominant_min_points) {
                rc = analyse_dominant(D, P, g_scalars, g_ptrs, vstride, stride, form,
                                      (size_t)S.table_W * S.n + (size_t)tl * S.table_W, &dom, &use_dom, &launches);
                if (rc) return rc;
            }
            plan_mults += P.field_mults();  // after the analysis: the statistics report the pairs really made
            plan_pairs += (double)P.m;
            rc = run_group(D, P, d_bases, d_inf, g_scalars, g_ptrs, vstride, stride, form, table_stride, val_offset, &launches, 3,
                           false, use_dom ? &dom : nullptr);
            if (rc) return rc;
            if (!D.finish_on_host && (rc = fetch_device_finish(D, g, D.stream))) return rc;
            COZK_CUDA(cudaEventRecord(D.ev[6], D.stream));
            COZK_CUDA(cudaStreamSynchronize(D.stream));
            add_stage_times(D);
            {
                auto h0 = std::chrono::steady_clock::now();
                if (D.finish_on_host) {
                    FinishArgs F{D.host_sums.data(), P.g, P.Wb, P.c, P.NS, P.log_l, pass_out + v0 * 72};
                    for (size_t v = 0; v < g; ++v) finish_body(v, F);
                } else {
                    host_normalize_batch(D.host_sums.data(), g, pass_out + v0 * 72);
                }
                host_finish_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
            }
        }
    }
    COZK_CUDA(cudaEventRecord(D.ev[7], D.stream));
    COZK_CUDA(cudaStreamSynchronize(D.stream));
    float total_ms = 0;
    cudaEventElapsedTime(&total_ms, D.ev[0], D.ev[7]);
    // ev[7] is recorded after the host-side finish, so total_ms (ev[0] -> ev[7]) already contains it
    D.stats[5] += host_finish_ms;
    D.stats[6] = total_ms;
    D.stats[0] = std::max(0.0, total_ms - (D.stats[1] + D.stats[2] + D.stats[3] + D.stats[4] + D.stats[5]));
    D.stats[7] = launches;
    D.stats[8] = last_c;
    D.stats[9] = last_W;
    D.stats[10] = plan_mults;
    D.stats[11] = plan_pairs;
    if (passes > 1) {
        std::vector<uint8_t> col(passes * 72);
        for (size_t v = 0; v < k; ++v) {
            for (size_t p = 0; p < passes; ++p) memcpy(col.data() + 72 * p, partial.data() + (p * k + v) * 72, 72);
            host_sum(col.data(), passes, out + 72 * v);
        }
    }
    return COZK_OK;
}

int msm_dispatch(cozk_ctx* ctx, int only_device, cozk_srs srs, size_t base_offset, size_t n,
                        const void* const* host_scalars, const void* const* dev_scalars, size_t k, size_t stride, int form,
                        unsigned max_bits, void* out) {
    if (!ctx || !out || k == 0 || (!host_scalars && !dev_scalars)) {
        set_error("null pointer or k == 0");
        return COZK_ERR_INVALID_ARG;
    }
    if ((form != COZK_MONT && form != COZK_CANON) || stride < 32 || (stride & 15)) {
        set_error("bad scalar form or stride (stride must be >= 32 and a multiple of 16)");
        return COZK_ERR_INVALID_ARG;
    }
    const void* const* sc = host_scalars ? host_scalars : dev_scalars;
    for (size_t j = 0; j < k; ++j) {
        if (!sc[j] && n) {
            set_error("null scalar vector");
            return COZK_ERR_INVALID_ARG;
        }
        if (dev_scalars && ((uintptr_t)sc[j] & 15)) {
            set_error("device scalar vectors must be 16-byte aligned");
            return COZK_ERR_INVALID_ARG;
        }
    }
    SrsEntry S;
    {
        int lrc = srs_lookup(ctx, srs, &S);
        if (lrc) return lrc;
    }
    if (base_offset > S.n || n > S.n - base_offset) {
        set_error("Key length error: base_offset + n exceeds the registered SRS");
        return COZK_ERR_KEY_LENGTH;
    }
    uint8_t* o = reinterpret_cast<uint8_t*>(out);
    if (n == 0) {
        for (size_t j = 0; j < k; ++j) {
            memset(o + 72 * j, 0, 72);
            o[72 * j + 64] = 1;
        }
        return COZK_OK;
    }
    if (!S.slices.empty()) {
        // Sliced SRS: every device holds only its point range (with its own table); the call is cut along the slices, the
        // devices run side by side (one host thread each) and the partial sums are added on the host - the reference's
        // split_ck + combine_comm scheme (co-noir-spartan/co-spartan/src/utils.rs:38-83, snarks-core/src/poly/commitment.rs:56-63).
        if (!host_scalars || only_device >= 0) {
            set_error("a sliced SRS takes host scalars through cozk_msm_batch");
            return COZK_ERR_INVALID_ARG;
        }
        struct Part {
            const SrsSlice* sl;
            size_t off, cnt, skip;  // offset inside the slice, points, points of the call before this part
        };
        std::vector<Part> parts;
        for (const SrsSlice& sl : S.slices) {
            const size_t a = std::max(base_offset, sl.lo), b = std::min(base_offset + n, sl.lo + sl.len);
            if (a < b) parts.push_back({&sl, a - sl.lo, b - a, a - base_offset});
        }
        const size_t np = parts.size();
        std::vector<int> rcs(np, COZK_OK);
        std::vector<std::string> errs(np);
        std::vector<uint8_t> partial(np * k * 72);
        std::vector<std::vector<const void*>> ptrs(np, std::vector<const void*>(k));
        auto work = [&](size_t pi) {
            const Part& pt = parts[pi];
            for (size_t j = 0; j < k; ++j) ptrs[pi][j] = reinterpret_cast<const uint8_t*>(host_scalars[j]) + pt.skip * stride;
            rcs[pi] = run_on_device(ctx, pt.sl->dev, *pt.sl->entry, pt.off, pt.cnt, ptrs[pi].data(), nullptr, k, stride, form, max_bits,
                                    &partial[pi * k * 72], np > 1 ? (long)ctx->opt_stream_min_points_sliced : -1);
            if (rcs[pi]) errs[pi] = g_error;
        };
        std::vector<std::thread> th;
        for (size_t pi = 1; pi < np; ++pi) th.emplace_back(work, pi);
        work(0);
        for (auto& t : th) t.join();
        for (size_t pi = 0; pi < np; ++pi) {
            if (rcs[pi]) {
                set_error(errs[pi]);
                return rcs[pi];
            }
        }
        std::vector<uint8_t> col(np * 72);
        for (size_t j = 0; j < k; ++j) {
            for (size_t pi = 0; pi < np; ++pi) memcpy(&col[72 * pi], &partial[(pi * k + j) * 72], 72);
            host_sum(col.data(), np, o + 72 * j);
        }
        return COZK_OK;
    }
    int nd = (int)ctx->devs.size();
    if (only_device >= 0 || nd == 1) {
        int d = only_device >= 0 ? only_device : 0;
        if (d >= nd) {
            set_error("device index out of range");
            return COZK_ERR_INVALID_ARG;
        }
        return run_on_device(ctx, d, S, base_offset, n, host_scalars, dev_scalars, k, stride, form, max_bits, o);
    }
    // several devices: shard by vector when there are enough of them, else by point range; combine on the host
    if (!host_scalars) {
        set_error("device-resident scalars need a device index (cozk_msm_batch_device)");
        return COZK_ERR_INVALID_ARG;
    }
    std::vector<int> rcs(nd, COZK_OK);
    std::vector<std::string> errs(nd);
    std::vector<std::thread> th;
    if (k >= (size_t)nd) {
        for (int d = 0; d < nd; ++d) {
            size_t v0 = k * d / nd, v1 = k * (d + 1) / nd;
            th.emplace_back([&, d, v0, v1] {
                rcs[d] = run_on_device(ctx, d, S, base_offset, n, host_scalars + v0, nullptr, v1 - v0, stride, form, max_bits,
                                       o + 72 * v0);
                if (rcs[d]) errs[d] = g_error;
            });
        }
        for (auto& t : th) t.join();
    } else {
        std::vector<uint8_t> partial((size_t)nd * k * 72);
        std::vector<std::vector<const void*>> ptrs(nd, std::vector<const void*>(k));
        std::vector<size_t> los(nd + 1);
        for (int d = 0; d <= nd; ++d) los[d] = n * (size_t)d / nd;
        for (int d = 0; d < nd; ++d) {
            for (size_t j = 0; j < k; ++j) ptrs[d][j] = reinterpret_cast<const uint8_t*>(host_scalars[j]) + los[d] * stride;
            th.emplace_back([&, d] {
                size_t cnt = los[d + 1] - los[d];
                if (cnt == 0) {
                    for (size_t j = 0; j < k; ++j) {
                        memset(&partial[((size_t)d * k + j) * 72], 0, 72);
                        partial[((size_t)d * k + j) * 72 + 64] = 1;
                    }
                    return;
                }
                rcs[d] = run_on_device(ctx, d, S, base_offset + los[d], cnt, ptrs[d].data(), nullptr, k, stride, form, max_bits,
                                       &partial[(size_t)d * k * 72]);
                if (rcs[d]) errs[d] = g_error;
            });
        }
        for (auto& t : th) t.join();
        std::vector<uint8_t> col((size_t)nd * 72);
        for (size_t j = 0; j < k; ++j) {
            for (int d = 0; d < nd; ++d) memcpy(&col[72 * d], &partial[((size_t)d * k + j) * 72], 72);
            host_sum(col.data(), nd, o + 72 * j);
        }
    }
    for (int d = 0; d < nd; ++d) {
        if (rcs[d]) {
            set_error(errs[d]);
            return rcs[d];
        }
    }
    return COZK_OK;
}

SrsMem::~SrsMem() {
    if (!owns) return;
    int cur = -1;
    cudaGetDevice(&cur);
    for (size_t j = 0; j < bases.size(); ++j) {
        if (!bases[j] && !inf[j]) continue;
        cudaSetDevice(cuda_id[j]);
        if (bases[j]) cudaFree(bases[j]);  // cudaFree waits for the device: nothing in flight still reads the memory
        if (inf[j]) cudaFree(inf[j]);
    }
    if (cur >= 0) cudaSetDevice(cur);
}

int srs_lookup(cozk_ctx* ctx, cozk_srs srs, SrsEntry* out) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->srs.find(srs);
    if (it == ctx->srs.end()) {
        set_error("unknown SRS handle");
        return COZK_ERR_BAD_HANDLE;
    }
    *out = it->second;  // shares the device memory: a concurrent cozk_srs_release cannot free it under this call
    return COZK_OK;
}

// rows of the precomputed table for an SRS of n points on the given devices under the context's memory policy
// (1 = bases only)
static void srs_table_shape(const cozk_ctx* ctx, const std::vector<int>& devices, size_t n, uint32_t* c, uint32_t* W,
                            uint32_t force_c = 0) {
    *c = 0;
    *W = 1;
    if (n < 1024 || ctx->opt_table_max_bytes <= 0) return;
    uint32_t tc = force_c ? force_c : ctx->opt_table_window ? (uint32_t)ctx->opt_table_window : choose_table_window(n);
    uint32_t tw = windows_for(254, tc);
    const double table_bytes = (double)tw * (double)n * sizeof(affine);
    if (table_bytes > (double)ctx->opt_table_max_bytes) return;
    // never more than half of what is free on a device now (sort buffers, buckets and resident polynomials need the rest)
    for (int di : devices) {
        size_t free_b = 0, total_b = 0;
        if (cudaSetDevice(ctx->devs[di]->id) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess &&
            table_bytes > 0.5 * (double)free_b)
            return;
    }
    if ((double)tw * (double)n >= 2147483647.0) return;  // table indices share 31 bits with the point index; all-ones is the skip mark
    *c = tc;
    *W = tw;
}

// The sums of every table row over the whole SRS and over its power-of-two prefixes (down to 1024 points), stored behind
// the table: what the dominant-digit mode of the decompose pass needs (msm_kernels.cuh).  ONE segmented sum per group of
// rows: the points of a row fall into `levels` blocks - [0, n >> (L-1)), then [n >> (L-j), n >> (L-j-1)) - whose keys are
// born sorted, the accumulate stage adds every block up, and the host forms the prefix sums over the blocks and
// normalises them (levels x rows points).  Setup-time work.
__global__ void __launch_bounds__(256) k_totals_pairs(size_t n, uint32_t levels, uint32_t row0, uint32_t rows, uint32_t* keys, uint32_t* vals) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)rows * n) return;
    const uint32_t r = (uint32_t)(t / n);
    const size_t i = t - (size_t)r * n;
    uint32_t b = 0;
    while (b + 1 < levels && i >= (n >> (levels - 1 - b))) ++b;
    keys[t] = r * levels + b;
    vals[t] = (uint32_t)((size_t)(row0 + r) * n + i);
}

static int srs_compute_totals(cozk_ctx* ctx, int di, SrsEntry& S) {
    S.total_levels = 0;
    S.total_ok = 0;
    const uint32_t levels = SrsEntry::totals_levels_for(S.n);
    if (levels == 0 || S.inf(di)) return COZK_OK;  // bases at infinity: the totals would have to leave them out; not worth a second code path
    Device& D = *ctx->devs[di];
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    const size_t n = S.n;
    const uint32_t rows_per_pass = (uint32_t)std::max<size_t>(1, std::min<size_t>(S.table_W, ((size_t)1 << 28) / n));
    const AccTuning acc_tuning{(size_t)D.sm_count * 512, 0, (int)ctx->opt_acc_chunk_up, 0};
    std::vector<xyzz> sums((size_t)S.table_W * levels);
    double launches = 0;
    int rc;
    for (uint32_t row0 = 0; row0 < S.table_W; row0 += rows_per_pass) {
        const uint32_t rows = std::min(rows_per_pass, S.table_W - row0);
        MsmPlan P;
        P.n = n;
        P.g = 1;
        P.acc = acc_tuning;
        P.total_buckets = (size_t)rows * levels;
        plan_set_pairs(P, (size_t)rows * n);
        if ((rc = D.keys_b.ensure(P.m * 4)) || (rc = D.vals_b.ensure(P.m * 4)) || (rc = D.buckets.ensure(P.total_buckets * sizeof(xyzz))))
            return rc;
        k_totals_pairs<<<grid_for(P.m, 256), 256, 0, D.stream>>>(n, levels, row0, rows, D.keys_b.as<uint32_t>(), D.vals_b.as<uint32_t>());
        COZK_CUDA(cudaGetLastError());
        if ((rc = run_accumulate(D, P, D.keys_b.as<uint32_t>(), D.vals_b.as<uint32_t>(), S.bases(di), D.buckets.as<xyzz>(), &launches)))
            return rc;
        COZK_CUDA(cudaMemcpyAsync(&sums[(size_t)row0 * levels], D.buckets.p, P.total_buckets * sizeof(xyzz), cudaMemcpyDeviceToHost, D.stream));
        COZK_CUDA(cudaStreamSynchronize(D.stream));
    }
    // total of level j (the first n >> j points) = blocks 0 .. levels-1-j; only exact halvings of the SRS length are used
    uint64_t ok = 0;
    std::vector<affine> host_totals((size_t)levels * S.table_W);
    memset(host_totals.data(), 0, host_totals.size() * sizeof(affine));
    for (uint32_t j = 0; j < levels; ++j)
        if (((n >> j) << j) == n) ok |= (uint64_t)1 << j;
    for (uint32_t r = 0; r < S.table_W; ++r) {
        xyzz acc = xyzz_identity();
        for (uint32_t b = 0; b < levels; ++b) {
            acc = xyzz_add(acc, sums[(size_t)r * levels + b]);
            const uint32_t j = levels - 1 - b;
            uint8_t wire[72];
            xyzz_to_wire(acc, wire);
            if (wire[64]) ok &= ~((uint64_t)1 << j);  // a sum that is the identity has no affine form: plain path for this length
            else memcpy(&host_totals[(size_t)j * S.table_W + r], wire, 64);
        }
    }
    COZK_CUDA(cudaMemcpyAsync(S.bases(di) + (size_t)S.table_W * n, host_totals.data(), host_totals.size() * sizeof(affine),
                              cudaMemcpyHostToDevice, D.stream));
    COZK_CUDA(cudaStreamSynchronize(D.stream));
    S.total_levels = levels;
    S.total_ok = ok;
    return COZK_OK;
}

// Where the bases of a registration come from: host memory, or the memory of one device of the context.
struct SrsSource {
    const void* host = nullptr;
    size_t stride = 64;
    const uint8_t* host_inf = nullptr;
    const void* dev = nullptr;
    int dev_index = -1;
    const uint8_t* dev_inf = nullptr;
};

// n points -> every device in `devices` gets the bases (+ flags), its precomputed table and the row totals.  The
// devices work side by side (one host thread each).
static int srs_register_core(cozk_ctx* ctx, const SrsSource& src, size_t n, const std::vector<int>& devices, SrsEntry* out,
                             uint32_t force_table_c = 0) {
    SrsEntry S;
    S.n = n;
    S.mem = std::make_shared<SrsMem>();
    const size_t nd = ctx->devs.size();
    S.mem->cuda_id.resize(nd);
    for (size_t d = 0; d < nd; ++d) S.mem->cuda_id[d] = ctx->devs[d]->id;
    S.mem->bases.assign(nd, nullptr);
    S.mem->inf.assign(nd, nullptr);
    srs_table_shape(ctx, devices, n, &S.table_c, &S.table_W, force_table_c);
    bool any_inf = false;
    if (src.host_inf) {
        for (size_t i = 0; i < n && !any_inf; ++i) any_inf = src.host_inf[i] != 0;
    } else if (src.dev_inf && n) {
        std::vector<uint8_t> flags(n);
        Device& S0 = *ctx->devs[src.dev_index];
        COZK_CUDA(cudaSetDevice(S0.id));
        COZK_CUDA(cudaMemcpyAsync(flags.data(), src.dev_inf, n, cudaMemcpyDeviceToHost, S0.stream));
        COZK_CUDA(cudaStreamSynchronize(S0.stream));
        for (size_t i = 0; i < n && !any_inf; ++i) any_inf = flags[i] != 0;
    }
    const size_t entries = std::max<size_t>(n, 1) * S.table_W + (size_t)S.table_W * (SrsEntry::totals_levels_for(n) + 1);
    std::vector<int> rcs(devices.size(), COZK_OK);
    std::vector<std::string> errs(devices.size());
    std::vector<SrsEntry> per_dev(devices.size(), S);  // each thread fills its own totals fields
    auto work = [&](size_t k) -> int {
        const int di = devices[k];
        Device& D = *ctx->devs[di];
        affine* d = nullptr;
        uint8_t* dinf = nullptr;
        {
            std::lock_guard<std::mutex> lock(D.mu);
            COZK_CUDA(cudaSetDevice(D.id));
            COZK_CUDA(cudaMalloc(&d, entries * sizeof(affine)));
            S.mem->bases[di] = d;  // from here on the memory is released with S.mem
            if (any_inf) {
                COZK_CUDA(cudaMalloc(&dinf, n));
                S.mem->inf[di] = dinf;
            }
            // Copies go through the engine's own stream and are synchronised there.  A plain cudaMemcpy from pageable memory
            // returns once the data is STAGED - its DMA runs on the legacy default stream, which the engine's non-blocking
            // streams do not wait for, so the table build could read row 0 before the tail of the copy had landed.
            if (src.host) {
                if (src.stride == 64) COZK_CUDA(cudaMemcpyAsync(d, src.host, n * 64, cudaMemcpyHostToDevice, D.stream));
                else COZK_CUDA(cudaMemcpy2DAsync(d, 64, src.host, src.stride, 64, n, cudaMemcpyHostToDevice, D.stream));
                if (any_inf) COZK_CUDA(cudaMemcpyAsync(dinf, src.host_inf, n, cudaMemcpyHostToDevice, D.stream));
            } else {
                const int sid = ctx->devs[src.dev_index]->id;
                COZK_CUDA(cudaMemcpyPeerAsync(d, D.id, src.dev, sid, n * sizeof(affine), D.stream));
                if (any_inf) COZK_CUDA(cudaMemcpyPeerAsync(dinf, D.id, src.dev_inf, sid, n, D.stream));
            }
            if (S.table_W > 1 && n) {
                // slabs of 2^20 points: the chain in XYZZ into a scratch array, then one inversion per point for all its rows
                // (table_body, one inversion per row, is the contract).  Registration of 2^20 / 2^22 / 2^24 points on one GPU:
                // 0.277 / 1.06 / 4.14 s row by row, 0.057 / 0.22 / 1.03 s this way
                const size_t slab = std::min<size_t>(n, (size_t)1 << 20);
                xyzz* tmp = nullptr;
                // (scratch from the device's stream-ordered pool: registrations in a row reuse it, nothing synchronises the device)
                if (S.table_W <= TABLE_MAX_ROWS && !ctx->opt_table_rowwise && n >= 8192 &&  // (tiny tables: not worth the scratch)
                    cudaMallocAsync(reinterpret_cast<void**>(&tmp), (size_t)(S.table_W - 1) * slab * sizeof(xyzz), D.stream) == cudaSuccess) {
                    for (size_t first = 0; first < n; first += slab) {
                        TableSlabArgs A{d, tmp, n, first, slab, S.table_c, S.table_W};
                        k_table_chain<<<grid_for(slab, 128), 128, 0, D.stream>>>(A);
                        k_table_norm<<<grid_for(slab, 128), 128, 0, D.stream>>>(A);
                    }
                    cudaError_t e = cudaGetLastError();
                    cudaFreeAsync(tmp, D.stream);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
                    if (e != cudaSuccess) {
                        set_error(std::string("SRS table build failed: ") + cudaGetErrorString(e));
                        return COZK_ERR_CUDA;
                    }
                } else {
                    cudaGetLastError();  // (a failed scratch allocation is not an error: fall back to the row-by-row kernel)
                    TableArgs A{d, n, S.table_c, S.table_W};
                    k_build_table<<<grid_for(n, 128), 128, 0, D.stream>>>(A);
                    COZK_CUDA(cudaGetLastError());
                }
            }
            COZK_CUDA(cudaStreamSynchronize(D.stream));
        }
        return srs_compute_totals(ctx, di, per_dev[k]);
    };
    auto guarded = [&](size_t k) {
        rcs[k] = work(k);
        if (rcs[k]) errs[k] = g_error;
    };
    std::vector<std::thread> th;
    for (size_t k = 1; k < devices.size(); ++k) th.emplace_back(guarded, k);
    if (!devices.empty()) guarded(0);
    for (auto& t : th) t.join();
    for (size_t k = 0; k < devices.size(); ++k) {
        if (rcs[k]) {
            set_error(errs[k].empty() ? "SRS registration failed" : errs[k]);
            return rcs[k];  // S.mem goes out of scope: everything allocated so far is given back
        }
    }
    if (!devices.empty()) {
        // the same points and the same plan on every device: the totals agree; keep the intersection of the usable levels
        S.total_levels = per_dev[0].total_levels;
        S.total_ok = per_dev[0].total_ok;
        for (size_t k = 1; k < devices.size(); ++k) {
            S.total_levels = std::min(S.total_levels, per_dev[k].total_levels);
            S.total_ok &= per_dev[k].total_ok;
        }
    }
    *out = S;
    return COZK_OK;
}

static std::vector<int> all_devices(const cozk_ctx* ctx) {
    std::vector<int> v(ctx->devs.size());
    for (size_t d = 0; d < v.size(); ++d) v[d] = (int)d;
    return v;
}
static cozk_srs srs_publish(cozk_ctx* ctx, const SrsEntry& S) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    const cozk_srs h = ctx->next_handle++;
    ctx->srs[h] = S;
    return h;
}

int srs_register_from_device(cozk_ctx* ctx, int device_index, const void* d_bases64, const uint8_t* d_inf, size_t n, cozk_srs* out,
                             int only_device, uint32_t force_table_c) {
    if (!ctx || !out || (!d_bases64 && n) || device_index < 0 || device_index >= (int)ctx->devs.size() ||
        only_device >= (int)ctx->devs.size()) {
        set_error("bad argument");
        return COZK_ERR_INVALID_ARG;
    }
    SrsSource src;
    src.dev = d_bases64;
    src.dev_index = device_index;
    src.dev_inf = d_inf;
    SrsEntry S;
    int rc = srs_register_core(ctx, src, n, only_device >= 0 ? std::vector<int>{only_device} : all_devices(ctx), &S, force_table_c);
    if (rc) return rc;
    *out = srs_publish(ctx, S);
    return COZK_OK;
}

int msm_ragged_device(cozk_ctx* ctx, int device, cozk_srs srs, const size_t* offsets, const size_t* lens,
                      const void* const* dev_scalars, size_t k, size_t stride, int form, void* out) {
    if (!ctx || !offsets || !lens || !dev_scalars || !out || k == 0 || k > 4096 || device < 0 || device >= (int)ctx->devs.size() ||
        stride < 32 || (stride & 15)) {
        set_error("ragged batch: bad argument");
        return COZK_ERR_INVALID_ARG;
    }
    SrsEntry S;
    int rc = srs_lookup(ctx, srs, &S);
    if (rc) return rc;
    if (!S.bases(device)) {
        set_error("ragged batch: the SRS does not live on this device");
        return COZK_ERR_INVALID_ARG;
    }
    size_t total = 0;
    std::vector<uint32_t> h_start(k + 1), h_base(k);
    for (size_t j = 0; j < k; ++j) {
        if (offsets[j] > S.n || lens[j] > S.n - offsets[j]) {
            set_error("Key length error: a vector of the ragged batch exceeds the registered SRS");
            return COZK_ERR_KEY_LENGTH;
        }
        if (!dev_scalars[j] || lens[j] == 0 || ((uintptr_t)dev_scalars[j] & 15)) {
            set_error("ragged batch: empty, null or misaligned vector");
            return COZK_ERR_INVALID_ARG;
        }
        h_start[j] = (uint32_t)total;
        h_base[j] = (uint32_t)offsets[j];
        total += lens[j];
    }
    h_start[k] = (uint32_t)total;
    const size_t max_buckets = (size_t)1 << 25;
    Device& D = *ctx->devs[device];
    const AccTuning acc_tuning{(size_t)D.sm_count * 512, (int)ctx->opt_acc_chunk, (int)ctx->opt_acc_chunk_up, (int)ctx->opt_group_l,
                               (int)ctx->opt_reduce_2d};
    MsmPlan P = make_plan((total + k - 1) / k, (uint32_t)k, 254, max_buckets, (uint32_t)ctx->opt_window, ctx->opt_window ? 0 : S.table_c,
                          acc_tuning);
    const bool table = !ctx->opt_window && S.table_c != 0;
    if (table && P.W > S.table_W) {
        set_error("ragged batch: the SRS table has too few rows");
        return COZK_ERR_INVALID_ARG;
    }
    if (P.total_buckets > max_buckets || (double)P.W * (double)total > (double)ctx->opt_group_pairs || total >= ((size_t)1 << 31)) {
        set_error("ragged batch: over the bucket / pair budget");
        return COZK_ERR_INVALID_ARG;
    }
    plan_set_pairs(P, (size_t)P.W * total);
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    D.sort_digit_bits = ctx->opt_sort_digit_bits;
    D.affine_rounds = ctx->opt_affine_rounds;
    D.affine_min_pairs = ctx->opt_affine_min_pairs;
    for (double& st : D.stats) st = 0;
    double launches = 0;
    if ((rc = D.rag.ensure((2 * k + 1) * sizeof(uint32_t))) || (rc = D.vec_ptrs.ensure(2 * 4096 * sizeof(void*)))) return rc;
    uint32_t* d_start = D.rag.as<uint32_t>();
    uint32_t* d_base = d_start + (k + 1);
    COZK_CUDA(cudaEventRecord(D.ev[0], D.stream));
    COZK_CUDA(cudaMemcpyAsync(d_start, h_start.data(), (k + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, D.stream));
    COZK_CUDA(cudaMemcpyAsync(d_base, h_base.data(), k * sizeof(uint32_t), cudaMemcpyHostToDevice, D.stream));
    COZK_CUDA(cudaMemcpyAsync(D.vec_ptrs.p, dev_scalars, k * sizeof(void*), cudaMemcpyHostToDevice, D.stream));
    COZK_CUDA(cudaStreamSynchronize(D.stream));  // the three host arrays are pageable
    DecomposeArgs rag = {};
    rag.rag_start = d_start;
    rag.rag_base = d_base;
    rag.total = total;
    rc = run_group(D, P, S.bases(device), S.inf(device), nullptr, D.vec_ptrs.as<const uint8_t*>(), 0, stride, form, table ? S.n : 0, 0,
                   &launches, 3, false, nullptr, &rag);
    if (rc) return rc;
    uint8_t* o = reinterpret_cast<uint8_t*>(out);
    if (!D.finish_on_host && (rc = fetch_device_finish(D, k, D.stream))) return rc;
    COZK_CUDA(cudaEventRecord(D.ev[6], D.stream));
    COZK_CUDA(cudaStreamSynchronize(D.stream));
    add_stage_times(D);
    if (D.finish_on_host) {
        FinishArgs F{D.host_sums.data(), P.g, P.Wb, P.c, P.NS, P.log_l, o};
        for (size_t v = 0; v < k; ++v) finish_body(v, F);
    } else {
        host_normalize_batch(D.host_sums.data(), k, o);
    }
    COZK_CUDA(cudaEventRecord(D.ev[7], D.stream));
    COZK_CUDA(cudaStreamSynchronize(D.stream));
    float total_ms = 0;
    cudaEventElapsedTime(&total_ms, D.ev[0], D.ev[7]);
    D.stats[6] = total_ms;
    D.stats[7] = launches;
    D.stats[8] = P.c;
    D.stats[9] = P.W;
    D.stats[10] = P.field_mults();
    D.stats[11] = (double)P.m;
    return COZK_OK;
}

}  // namespace cozk

using namespace cozk;

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* cozk_last_error(void) { return g_error.c_str(); }

int cozk_init(cozk_ctx** out, const int* device_ids, int n_devices) {
    if (!out) {
        set_error("null out pointer");
        return COZK_ERR_INVALID_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
        return COZK_ERR_NO_DEVICE;
    }
    if (n_devices <= 0) n_devices = 1;
    std::unique_ptr<cozk_ctx> ctx(new cozk_ctx);
    for (int i = 0; i < n_devices; ++i) {
        int id = device_ids ? device_ids[i] : i;
        if (id < 0 || id >= count) {
            set_error("device id out of range");
            return COZK_ERR_NO_DEVICE;
        }
        std::unique_ptr<Device> D(new Device);
        D->id = id;
        COZK_CUDA(cudaSetDevice(id));
        cudaDeviceProp prop;
        COZK_CUDA(cudaGetDeviceProperties(&prop, id));
        if (prop.major < 10) {
            set_error("this library is built for sm_100a (B200) only");
            return COZK_ERR_NO_DEVICE;
        }
        D->sm_count = prop.multiProcessorCount;
        {
            // keep freed stream-ordered allocations (device-resident polynomials, rep3poly.cu) in the pool
            cudaMemPool_t pool;
            COZK_CUDA(cudaDeviceGetDefaultMemPool(&pool, id));
            uint64_t keep = UINT64_MAX;
            COZK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        {
            int src = sort_setup_device();
            if (src) return src;
            if ((src = affine_setup_device())) return src;
        }
        COZK_CUDA(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking));
        COZK_CUDA(cudaStreamCreateWithFlags(&D->copy_stream, cudaStreamNonBlocking));
        for (auto& ev : D->ev) COZK_CUDA(cudaEventCreate(&ev));
        for (auto& ev : D->copy_done) COZK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->devs.push_back(std::move(D));
    }
    *out = ctx.release();
    return COZK_OK;
}

void cozk_destroy(cozk_ctx* ctx) {
    if (!ctx) return;
    ctx->polys.clear();  // published entries own their memory (PolyEntry::hold): stream-ordered frees, drained below
    for (auto& D : ctx->devs) {
        cudaMemPool_t pool;
        if (cudaSetDevice(D->id) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, D->id) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
    }
    ctx->srs.clear();  // the entries own their device memory (SrsMem)
    delete ctx;
}

int cozk_device_count(const cozk_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int cozk_srs_register(cozk_ctx* ctx, const void* bases, size_t n, size_t stride_bytes, const uint8_t* infinity, cozk_srs* out) {
    if (!ctx || !out || (!bases && n) || stride_bytes < 64) {
        set_error("null pointer or stride < 64");
        return COZK_ERR_INVALID_ARG;
    }
    SrsSource src;
    src.host = bases;
    src.stride = stride_bytes;
    src.host_inf = infinity;
    SrsEntry S;
    int rc = srs_register_core(ctx, src, n, all_devices(ctx), &S);
    if (rc) return rc;
    *out = srs_publish(ctx, S);
    return COZK_OK;
}

// Device d of the context gets points [n*d/D, n*(d+1)/D) only - with their own table and row totals - instead of a copy
// of everything: SURVEY.md 8(e) "each GPU holds only its slice", the reference's split_ck
// (co-noir-spartan/co-spartan/src/utils.rs:38-83).  Every MSM over such an SRS is sharded by point range along the slices.
int cozk_srs_register_sliced(cozk_ctx* ctx, const void* bases, size_t n, size_t stride_bytes, const uint8_t* infinity, cozk_srs* out) {
    if (!ctx || !out || (!bases && n) || stride_bytes < 64) {
        set_error("null pointer or stride < 64");
        return COZK_ERR_INVALID_ARG;
    }
    const size_t nd = ctx->devs.size();
    SrsEntry P;
    P.n = n;
    std::vector<SrsSlice> slices;
    for (size_t d = 0; d < nd; ++d) {
        SrsSlice sl;
        sl.dev = (int)d;
        sl.lo = n * d / nd;
        sl.len = n * (d + 1) / nd - sl.lo;
        if (sl.len) slices.push_back(sl);
    }
    std::vector<int> rcs(slices.size(), COZK_OK);
    std::vector<std::string> errs(slices.size());
    auto work = [&](size_t k) {
        SrsSlice& sl = slices[k];
        SrsSource src;
        src.host = reinterpret_cast<const uint8_t*>(bases) + sl.lo * stride_bytes;
        src.stride = stride_bytes;
        src.host_inf = infinity ? infinity + sl.lo : nullptr;
        sl.entry = std::make_shared<SrsEntry>();
        rcs[k] = srs_register_core(ctx, src, sl.len, std::vector<int>{sl.dev}, sl.entry.get());
        if (rcs[k]) errs[k] = g_error;
    };
    std::vector<std::thread> th;
    for (size_t k = 1; k < slices.size(); ++k) th.emplace_back(work, k);
    if (!slices.empty()) work(0);
    for (auto& t : th) t.join();
    for (size_t k = 0; k < slices.size(); ++k) {
        if (rcs[k]) {
            set_error(errs[k]);
            return rcs[k];
        }
    }
    if (slices.empty()) {  // n == 0: an ordinary empty SRS
        SrsSource src;
        src.host = bases;
        src.stride = stride_bytes;
        int rc = srs_register_core(ctx, src, 0, all_devices(ctx), &P);
        if (rc) return rc;
    } else {
        P.slices = slices;
    }
    *out = srs_publish(ctx, P);
    return COZK_OK;
}

int cozk_srs_register_device(cozk_ctx* ctx, int device_index, const void* d_bases64, size_t n, cozk_srs* out) {
    return srs_register_from_device(ctx, device_index, d_bases64, nullptr, n, out);
}

// Drops the handle.  Calls that looked the SRS up before this point keep its device memory alive until they return.
int cozk_srs_release(cozk_ctx* ctx, cozk_srs srs) {
    if (!ctx) return COZK_ERR_INVALID_ARG;
    SrsEntry S;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        auto it = ctx->srs.find(srs);
        if (it == ctx->srs.end()) {
            set_error("unknown SRS handle");
            return COZK_ERR_BAD_HANDLE;
        }
        S = it->second;
        ctx->srs.erase(it);
    }
    return COZK_OK;  // S leaves scope here, outside the context lock: the last reference frees the memory
}

int cozk_srs_len(cozk_ctx* ctx, cozk_srs srs, size_t* out_n) {
    if (!ctx || !out_n) return COZK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->srs.find(srs);
    if (it == ctx->srs.end()) {
        set_error("unknown SRS handle");
        return COZK_ERR_BAD_HANDLE;
    }
    *out_n = it->second.n;
    return COZK_OK;
}

int cozk_msm_batch(cozk_ctx* ctx, cozk_srs srs, size_t base_offset, size_t n, const void* const* scalars, size_t k,
                   size_t stride_bytes, int form, unsigned max_num_bits, void* out) {
    if (!scalars) {
        set_error("null scalars");
        return COZK_ERR_INVALID_ARG;
    }
    return msm_dispatch(ctx, -1, srs, base_offset, n, scalars, nullptr, k, stride_bytes, form, max_num_bits, out);
}

int cozk_msm_batch_device(cozk_ctx* ctx, int device_index, cozk_srs srs, size_t base_offset, size_t n,
                          const void* const* d_scalars, size_t k, size_t stride_bytes, int form, unsigned max_num_bits,
                          void* out) {
    if (!d_scalars || device_index < 0) {
        set_error("null scalars or negative device index");
        return COZK_ERR_INVALID_ARG;
    }
    return msm_dispatch(ctx, device_index, srs, base_offset, n, nullptr, d_scalars, k, stride_bytes, form, max_num_bits, out);
}

int cozk_msm_ragged_device(cozk_ctx* ctx, int device_index, cozk_srs srs, const size_t* base_offsets, const size_t* lens,
                           const void* const* d_scalars, size_t k, size_t stride_bytes, int form, void* out) {
    if (form != COZK_MONT && form != COZK_CANON) {
        set_error("bad scalar form");
        return COZK_ERR_INVALID_ARG;
    }
    return msm_ragged_device(ctx, device_index, srs, base_offsets, lens, d_scalars, k, stride_bytes, form, out);
}

int cozk_g1_sum(const void* points72, size_t count, void* out72) {
    if ((!points72 && count) || !out72) {
        set_error("null pointer");
        return COZK_ERR_INVALID_ARG;
    }
    host_sum(reinterpret_cast<const uint8_t*>(points72), count, reinterpret_cast<uint8_t*>(out72));
    return COZK_OK;
}

int cozk_set_option(cozk_ctx* ctx, const char* name, long value) {
    if (!ctx || !name) return COZK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!strcmp(name, "window")) {
        if (value != 0 && (value < (long)C_MIN || value > (long)C_MAX)) {
            set_error("window out of range");
            return COZK_ERR_INVALID_ARG;
        }
        ctx->opt_window = value;
    } else if (!strcmp(name, "acc_chunk")) {
        // pairs per thread at level 1 of the accumulate stage: 0 = choose per call (fill the last wave of threads)
        if (value != 0 && (value < 4 || value > 256)) return COZK_ERR_INVALID_ARG;
        ctx->opt_acc_chunk = value;
    } else if (!strcmp(name, "acc_chunk_up")) {
        // partial slots per thread at the serial accumulate levels >= 2 (0 = 32); never ACC_TILE, which selects the scan kernel
        if (value != 0 && (value < 2 || value > 128 || value == ACC_TILE)) return COZK_ERR_INVALID_ARG;
        ctx->opt_acc_chunk_up = value;
    } else if (!strcmp(name, "bulk_copy")) {
        if (value != 0 && value != 1) return COZK_ERR_INVALID_ARG;
        ctx->opt_bulk_copy = value;
    } else if (!strcmp(name, "chi_waves")) {
        if (value < 1 || value > 16) return COZK_ERR_INVALID_ARG;
        ctx->opt_chi_waves = value;
    } else if (!strcmp(name, "max_points_per_pass")) {
        if (value < 1 || value > (long)MAX_POINTS_PER_PASS) return COZK_ERR_INVALID_ARG;
        ctx->opt_max_points_per_pass = value;
    } else if (!strcmp(name, "sort_digit_bits")) {
        if (value < 7 || value > 11) return COZK_ERR_INVALID_ARG;  // 8 is the measured optimum (profiles/round2_summary.md)
        ctx->opt_sort_digit_bits = value;
    } else if (!strcmp(name, "group_l")) {
        // buckets per thread in the group step of the bucket reduce: a power of two, 0 = chosen from the bucket count
        if (value != 0 && (value < 1 || value > 64 || (value & (value - 1)))) return COZK_ERR_INVALID_ARG;
        ctx->opt_group_l = value;
    } else if (!strcmp(name, "dominant")) {
        // 1: whole-SRS calls look for windows dominated by one digit and use the row totals; 0: always the plain layout
        if (value != 0 && value != 1) return COZK_ERR_INVALID_ARG;
        ctx->opt_dominant = value;
    } else if (!strcmp(name, "dominant_min_points")) {
        if (value < 0) return COZK_ERR_INVALID_ARG;
        ctx->opt_dominant_min_points = value;
    } else if (!strcmp(name, "peer_direct")) {
        // multi-device linear combination: 1 = the summing kernel reads remote partials over peer mappings, 0 = staged copies
        if (value != 0 && value != 1) return COZK_ERR_INVALID_ARG;
        ctx->opt_peer_direct = value;
    } else if (!strcmp(name, "group_pairs")) {
        if (value < 1) return COZK_ERR_INVALID_ARG;
        ctx->opt_group_pairs = value;
    } else if (!strcmp(name, "stream_min_points")) {
        // a single host-resident vector of at least this many points is streamed in chunks (0 = never)
        if (value < 0) return COZK_ERR_INVALID_ARG;
        ctx->opt_stream_min_points = value;
    } else if (!strcmp(name, "table_rowwise")) {
        // 1: SRS tables are built by the row-by-row kernel (one inversion per row and point) instead of chain + batch normalisation
        if (value != 0 && value != 1) return COZK_ERR_INVALID_ARG;
        ctx->opt_table_rowwise = value;
    } else if (!strcmp(name, "open_one_batch_max_nv")) {
        if (value < 0 || value > 30) return COZK_ERR_INVALID_ARG;
        ctx->opt_open_one_batch_max_nv = value;
    } else if (!strcmp(name, "open_small_window")) {
        if (value != 0 && (value < (long)C_MIN || value > (long)C_MAX)) return COZK_ERR_INVALID_ARG;
        ctx->opt_open_small_window = value;
    } else if (!strcmp(name, "open_small_ragged")) {
        if (value != 0 && value != 1) return COZK_ERR_INVALID_ARG;
        ctx->opt_open_small_ragged = value;
    } else if (!strcmp(name, "reduce_2d")) {
        // bucket reduce: 1 = row / column form (two tree sums), 0 = group running sums + masked sums
        if (value != 0 && value != 1) return COZK_ERR_INVALID_ARG;
        ctx->opt_reduce_2d = value;
    } else if (!strcmp(name, "affine_rounds")) {
        // halving rounds of the batched-affine pre-reduction in front of the accumulate levels (0 = off)
        if (value < 0 || value > AFF_MAX_ROUNDS) return COZK_ERR_INVALID_ARG;
        ctx->opt_affine_rounds = value;
    } else if (!strcmp(name, "affine_min_pairs")) {
        if (value < 0) return COZK_ERR_INVALID_ARG;
        ctx->opt_affine_min_pairs = value;
    } else if (!strcmp(name, "chunk_min_points")) {
        if (value < 0) return COZK_ERR_INVALID_ARG;
        ctx->opt_chunk_min_points = value;
    } else if (!strcmp(name, "stream_min_points_sliced")) {
        if (value < 0) return COZK_ERR_INVALID_ARG;
        ctx->opt_stream_min_points_sliced = value;
    } else if (!strcmp(name, "stream_first_pct")) {
        if (value < 1 || value > 6400) return COZK_ERR_INVALID_ARG;
        ctx->opt_stream_first_pct = value;
    } else if (!strcmp(name, "stream_chunks")) {
        if (value < 0 || value > 64) return COZK_ERR_INVALID_ARG;  // 0 = chosen from the vector length
        ctx->opt_stream_chunks = value;
    } else if (!strcmp(name, "table_window")) {
        // window size of the tables of SRS registered from now on (0 = choose from the SRS length)
        if (value != 0 && (value < (long)C_MIN || value > (long)C_MAX)) return COZK_ERR_INVALID_ARG;
        ctx->opt_table_window = value;
    } else if (!strcmp(name, "open_small_log2")) {
        // opening keys created from now on: levels with at most 2^value quotient scalars share one batched MSM (0 = none)
        if (value < 0 || value > 20) return COZK_ERR_INVALID_ARG;
        ctx->opt_open_small_log2 = value;
    } else if (!strcmp(name, "table_max_mib")) {
        // memory an SRS registered from now on may spend on its 2^(c*w) * P table; 0 = no tables
        if (value < 0) return COZK_ERR_INVALID_ARG;
        ctx->opt_table_max_bytes = value << 20;
    } else {
        set_error("unknown option");
        return COZK_ERR_INVALID_ARG;
    }
    return COZK_OK;
}

int cozk_last_stats(cozk_ctx* ctx, double* out12) { return cozk_last_stats_device(ctx, 0, out12); }

int cozk_last_stats_device(cozk_ctx* ctx, int device_index, double* out12) {
    if (!ctx || !out12 || device_index < 0 || device_index >= (int)ctx->devs.size()) return COZK_ERR_INVALID_ARG;
    Device& D = *ctx->devs[device_index];
    std::lock_guard<std::mutex> lock(D.mu);
    for (int i = 0; i < 12; ++i) out12[i] = D.stats[i];
    return COZK_OK;
}

}  // extern "C"
