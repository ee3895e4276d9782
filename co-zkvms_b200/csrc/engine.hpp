// Internal definitions shared by the translation units of libcozk_msm.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cozk_msm.h"
#include "curve.cuh"

namespace cozk {

void set_error(const std::string& msg);

#define COZK_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            cozk::set_error(buf__);                                                                  \
            return COZK_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

// grow-only device scratch buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return COZK_OK;
        if (p) COZK_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        COZK_CUDA(cudaMalloc(&p, want));
        cap = want;
        return COZK_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct Device {
    int id = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // chunked calls: the upper accumulate levels + bucket merge of chunk i run here, beside level 1 of chunk i + 1
    cudaStream_t prep_stream = nullptr;  // chunked calls: decompose + sort of chunk i + 1 (high priority), beside level 1 of chunk i
    cudaStream_t l1b_stream = nullptr;   // chunked calls: level 1 of the odd chunks (the even ones run on `stream`): the tail of one chunk's kernel
                                         // overlaps the head of the next one's
    std::vector<cudaEvent_t> chunk_ev;   // chunked calls: 6 events per chunk, grown on demand
    std::mutex mu;  // one MSM at a time per device
    long affine_rounds = 0, affine_min_pairs = 0;  // options "affine_rounds" / "affine_min_pairs", copied from the context by every call
    long sort_digit_bits = 8;  // digit bits per sort pass on this device (option "sort_digit_bits", copied from the context by every call)
    DevBuf aff[2];  // batched-affine pre-reduction (affine.cu): round outputs, overflow lists, batch products (one set per chunk parity)
    DevBuf rag;  // ragged groups: start / base arrays of the vectors
    DevBuf scalars[2], vec_ptrs, keys_a, vals_a, keys_b, vals_b, sort_tmp,
           keys_a2, vals_a2, keys_b2, vals_b2,  // second set of sort buffers (chunked calls: chunk i + 1 is sorted while level 1 reads chunk i)
           buckets, buckets2[2], pk[4], pp[4], rs[2], rw[2], out, flush;
    std::mutex open_mu;  // one PST13 opening at a time per device: it owns the four buffers below across its MSM calls
    DevBuf open_in, open_r[2], open_q, open_qs;
    // dominant-digit analysis (msm_kernels.cuh, DomArgs): per-segment candidate digits, counters, modes, offsets, cursors
    DevBuf dom_cand, dom_counts, dom_mode, dom_off, dom_len, dom_cursor;
    std::vector<int32_t> h_dom_cand;
    std::vector<uint32_t> h_dom_counts, h_dom_mode;
    std::vector<uint64_t> h_dom_off, h_dom_len;
    cudaEvent_t ev[8] = {};
    cudaEvent_t copy_done[2] = {};
    double stats[12] = {};
    std::vector<xyzz> host_sums;  // per-window sums brought back for the host-side finish
    bool finish_on_host = false;
    ~Device();
};

// Device memory of one registered SRS.  Shared by the handle table and by every call in flight that looked the SRS up:
// cozk_srs_release only drops the table's reference, the memory goes when the last user lets go.
struct SrsMem {
    std::vector<int> cuda_id;        // CUDA device id per context device
    std::vector<affine*> bases;      // per context device: table_W rows of n points (row 0 = the bases) + row totals; null = not on this device
    std::vector<uint8_t*> inf;       // per context device, may be null
    bool owns = true;                // false: a view into another SRS's memory (row views of the totals computation)
    ~SrsMem();
};

struct SrsEntry;
// One slice of a sliced SRS (cozk_srs_register_sliced): points [lo, lo + len) live on context device `dev` only, as a
// complete single-device SRS with its own table and row totals.
struct SrsSlice {
    int dev = 0;
    size_t lo = 0, len = 0;
    std::shared_ptr<SrsEntry> entry;
};

struct SrsEntry {
    size_t n = 0;
    std::shared_ptr<SrsMem> mem;     // null for a sliced SRS (its slices own the memory)
    std::vector<SrsSlice> slices;    // non-empty: sliced SRS, every call is sharded by point range along these slices
    affine* bases(int dev) const { return mem && dev < (int)mem->bases.size() ? mem->bases[dev] : nullptr; }
    uint8_t* inf(int dev) const { return mem && dev < (int)mem->inf.size() ? mem->inf[dev] : nullptr; }
    uint32_t table_c = 0;            // window size the table rows were built for; 0 = no table
    uint32_t table_W = 1;            // rows: row w = 2^(table_c*w) * bases
    // Dominant-digit mode: behind the table, entry table_W * n + j * table_W + w holds the sum of the first n >> j points of
    // row w, for j = 0 .. total_levels - 1 (n >> j >= 1024): whole-SRS calls and power-of-two prefixes (a 2^16 polynomial
    // against a 2^22 SRS).  Bit j of total_ok: level j is usable (every row sum is a finite point).
    uint32_t total_levels = 0;
    uint64_t total_ok = 0;
    static uint32_t totals_levels_for(size_t n) {
        uint32_t l = 0;
        while (l < 40 && (n >> l) >= 1024) ++l;
        return l;
    }
};

// A device-resident polynomial (include/cozk_rep3.h).  d_data holds `total` coefficients; the handle covers
// [lo, lo + len) of them (Rep3DensePolynomial::chunk_range).
struct PolyEntry {
    int dev = 0;            // index into cozk_ctx::devs
    uint32_t kind = 0;      // POLY_SHARED / POLY_MONT / POLY_CANON (rep3_kernels.cuh)
    int user_kind = 0;      // the COZK_POLY_* constant it was created as
    unsigned bits = 0;      // max_num_bits hint for the MSM (8/16/32/64 for small unsigned kinds, 0 otherwise)
    uint8_t* d_data = nullptr;
    // Set when the entry is published in the handle table: owns d_data (stream-ordered free on the owning device's stream
    // when the last copy of the entry goes).  A call that looked the polynomial up holds a copy, so cozk_poly_release on
    // another thread cannot free the memory under it.
    std::shared_ptr<void> hold;
    size_t total = 0, lo = 0, len = 0;
    size_t elem_bytes() const { return kind == 0 ? 64 : 32; }
    const uint8_t* chunk() const { return d_data + lo * elem_bytes(); }
};

// Setup-time data of the PST13 opening (include/cozk_rep3.h): pair sums of every SRS level, the small levels
// concatenated into one SRS so that a single batched MSM serves all of them.
struct OpenKey {
    size_t nv = 0;
    std::vector<cozk_srs> level_srs;   // as given by the caller (not owned)
    cozk_srs big_srs = 0;              // owned: pair sums of the levels i < first_small, level after level, ONE SRS with one table
    std::vector<size_t> big_off;       // offset of level i inside big_srs
    size_t first_small = 0;            // levels [first_small, nv) go through one batched MSM
    cozk_srs small_srs = 0;            // owned: pair sums of the small levels, level after level
    size_t small_n = 0;
    std::vector<size_t> small_off;     // offset of level first_small + j inside small_srs
};

}  // namespace cozk

struct cozk_ctx;
namespace cozk {
// Register n points that already live on device `device_index` (d_inf: optional per-point infinity flags, device memory).
// only_device < 0: the SRS is replicated on every device of the context; >= 0: it lives on that device alone (the
// internal SRSs of an opening key, which only device 0 ever reads).
int srs_register_from_device(cozk_ctx* ctx, int device_index, const void* d_bases64, const uint8_t* d_inf, size_t n, cozk_srs* out,
                             int only_device = -1, uint32_t force_table_c = 0);
int srs_lookup(cozk_ctx* ctx, cozk_srs srs, SrsEntry* out);
// Where the 2^nv evaluations of an opening come from: host memory (staged through a scratch buffer) or device 0.
struct OpenSource {
    const void* host = nullptr;      // element i at host + i * stride
    const uint8_t* dev = nullptr;    // element i at dev + i * stride (share a of an AoS share array: stride 64)
    size_t stride = 32;
    int canon = 0;                   // device source holds canonical integers (small-scalar public polynomial)
};
// PST13 opening on device 0.  key == nullptr: the reference's schedule (one MSM per level over duplicated scalars);
// else pair sums + one batched MSM for the small levels.
int pst13_open_device(cozk_ctx* ctx, const cozk_srs* level_srs, const OpenKey* key, size_t nv, const OpenSource& src,
                      const void* point, void* out_proofs, void* out_eval);
// S[b] = P[2b] + P[2b+1] of a registered SRS into caller-provided device buffers on device 0 (half = n / 2 entries each).
int srs_pair_sums_into(cozk_ctx* ctx, cozk_srs srs, affine* d_out, uint8_t* d_inf, size_t* half_out);
int open_key_lookup(cozk_ctx* ctx, uint64_t h, OpenKey* out);
// Pair sort (sort.cu / sort_kernels.cuh): see there.
struct DecomposeArgs;
int sort_setup_device();
int sort_pairs(Device& D, cudaStream_t st, const DecomposeArgs* fused, size_t m, uint32_t key_bits, uint32_t** keys_out,
               uint32_t** vals_out, double* launches, cudaEvent_t after_first, int set = 0);
// Batched-affine pre-reduction of a sorted pair list (affine.cu / affine_kernels.cuh): see there.
constexpr int AFF_MAX_ROUNDS = 6;
struct AffineRoundArgs;
struct OvfAddArgsPub {  // layout of OvfAddArgs (affine_kernels.cuh), so that this header needs no kernel definitions
    const uint32_t* count;
    const uint32_t* keys;
    const affine* pts;
    xyzz* buckets;
    uint32_t cap;
};
struct AffineStage {
    const uint32_t* keys = nullptr;  // the reduced list: keys, vals (index into pts, or the skip mark), points
    const uint32_t* vals = nullptr;
    const affine* pts = nullptr;
    size_t m = 0;
    int rounds = 0;
    OvfAddArgsPub ovf[AFF_MAX_ROUNDS] = {};
};
int affine_setup_device();
int affine_reduce(Device& D, cudaStream_t st, int parity, size_t m, size_t total_buckets, const uint32_t* keys, const uint32_t* vals,
                  const affine* bases, int rounds, AffineStage* out, double* launches);
int affine_overflow_adds(cudaStream_t st, const AffineStage& S, xyzz* buckets, double* launches);
int affine_round_reference(cudaStream_t st, const AffineRoundArgs& A);
// the plain decompose kernel (msm.cu), for the test entry points of aux.cu
int launch_decompose(const DecomposeArgs& A, cudaStream_t st);
// Ragged batch on ONE device (msm.cu): vector j has lens[j] device-resident scalars (dense stride) and multiplies bases
// [offsets[j], offsets[j] + lens[j]) of the SRS; one decompose / sort / accumulate / reduce for all of them.  Returns
// COZK_ERR_INVALID_ARG (and touches nothing) when the group does not fit the bucket / pair budget: the caller falls back
// to one call per vector.
int msm_ragged_device(cozk_ctx* ctx, int device, cozk_srs srs, const size_t* offsets, const size_t* lens,
                      const void* const* dev_scalars, size_t k, size_t stride, int form, void* out);
// The one entry every public MSM call funnels into (msm.cu).  only_device < 0: use all devices of the context.
int msm_dispatch(cozk_ctx* ctx, int only_device, cozk_srs srs, size_t base_offset, size_t n, const void* const* host_scalars,
                 const void* const* dev_scalars, size_t k, size_t stride, int form, unsigned max_bits, void* out);
}  // namespace cozk

struct cozk_ctx {
    std::vector<std::unique_ptr<cozk::Device>> devs;
    std::mutex mu;  // guards the handle tables; the options below are atomics (set by cozk_set_option, read by calls in flight)
    std::map<uint64_t, cozk::SrsEntry> srs;
    std::map<uint64_t, cozk::PolyEntry> polys;
    std::map<uint64_t, cozk::OpenKey> open_keys;
    std::atomic<long> opt_open_small_log2 = 15;   // opening levels with at most 2^this quotient values share one batched MSM (measured at nv = 22 / 18: 13: 17.5 ms, 14: 17.2 / 4.13, 15: 17.0 / 3.91, 16: 17.5 / 4.42, 17: 18.0)
    std::atomic<long> opt_open_one_batch_max_nv = 20;  // opening keys up to this nv put ALL levels into the one ragged batch of "small" levels
    std::atomic<long> opt_open_small_window = 0;  // table window of the small levels' SRS of opening keys created from now on; 0 = cost model
    std::atomic<long> opt_open_small_ragged = 1;  // 1: the small levels of a keyed opening run as a ragged batch (their real scalars only); 0: zero-padded batch
    double rep3_stats[8] = {};
    uint64_t next_handle = 1;
    std::atomic<long> opt_dominant = 1;           // 1: whole-SRS calls look for windows dominated by one digit (constant co-jolt shares) and use the row totals
    std::atomic<long> opt_dominant_min_points = 1L << 21;  // ... when the call has at least this many (vector, point) pairs: the look costs ~35 us
    std::atomic<long> opt_peer_direct = 1;        // 1: kernels read other devices' partial results through peer mappings; 0: stage peer copies first
    std::atomic<long> opt_acc_chunk = 0;          // pairs per level-1 accumulate thread; 0 = chosen per call (msm_plan.hpp, choose_acc_l)
    std::atomic<long> opt_acc_chunk_up = 0;       // partial slots per thread at the serial accumulate levels >= 2; 0 = ACC_L
    std::atomic<long> opt_bulk_copy = 0;          // 1: ingest / chi kernels stream their slabs through shared memory with bulk copies (TMA); 0: plain per-lane
                                     // loads.  Measured equal or slower (2^22: ingest 0.161 against 0.152 ms, chi 2.21 against 2.20 ms; 2^20: chi
                                     // 0.67 against 0.59 ms): the access pattern is not what holds these kernels back.  Kept as an option.
    std::atomic<long> opt_chi_waves = 1;          // threads per polynomial of the chi kernels: enough for this many full waves of the device
    std::atomic<long> opt_max_points_per_pass = 1L << 26;  // longer calls run as several passes whose results are added on the host
    std::atomic<long> opt_sort_digit_bits = 8;   // digit bits per pass of the pair sort (7 .. 11)
    std::atomic<long> opt_group_l = 0;            // buckets per thread in the group step of the bucket reduce; 0 = chosen from the bucket count
    std::atomic<long> opt_window = 0;             // 0 = choose per call
    std::atomic<long> opt_group_pairs = 1L << 29; // (key, val) pairs per vector group (8 GiB of sort buffers; B200 has 180 GB)
    std::atomic<long> opt_stream_min_points = 1L << 20;  // host-resident single vectors this long are streamed in chunks (0 = never).  Pipelined on three
                                                         // streams (msm.cu): 2^20 3.78 -> 3.70 ms in 2 chunks (3.84 / 3.98 in 3 / 4), 2^21 7.02 -> 6.80, 2^22 12.3 -> 12.15
    std::atomic<long> opt_stream_min_points_sliced = 1L << 20;  // the same threshold for the parts of a call over a sliced SRS: several devices
                                                                // copy from one host buffer at once, every copy is slower, overlap pays earlier
    std::atomic<long> opt_reduce_2d = 0;                  // bucket reduce in row / column form (msm_plan.hpp, MsmPlan::reduce_2d)
    std::atomic<long> opt_affine_rounds = 0;              // batched-affine pre-reduction rounds in front of the accumulate levels (affine_kernels.cuh); 0 = off
    std::atomic<long> opt_affine_min_pairs = 1L << 20;    // ... for pair lists of at least this many entries
    std::atomic<long> opt_chunk_min_points = 0;           // device-resident single vectors this long run in chunks too (0 = never, the default).  Measured: there
                                                          // is no copy to hide, and what runs beside the level-1 kernels (the next chunk's sort, the upper levels)
                                                          // is throughput work itself: 2^20 3.16 ms in one piece, 3.29 / 3.42 / 3.53 in 2 / 3 / 4 chunks; 2^24 37.3, 38.0 / 38.8 / 39.5
    std::atomic<long> opt_stream_first_pct = 50;          // chunked calls: length of the first chunk in percent of an equal share
    std::atomic<long> opt_stream_chunks = 0;              // 0 = auto: 2 chunks (msm.cu has the measurements)
    std::atomic<long> opt_table_rowwise = 0;            // 1: build SRS tables with the row-by-row kernel (the contract form); 0: chain + one inversion per point
    std::atomic<long> opt_table_window = 0;             // 0 = choose_table_window(n) at registration
    std::atomic<long> opt_table_max_bytes = 64L << 30;  // per-SRS budget for the precomputed 2^(c*w) * P table (12 x 4 GiB at 2^26; the GPU has 180 GB), and never more than half of the free device memory; 0 disables tables
};
