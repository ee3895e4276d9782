/* cozk_msm.h - C ABI of the B200-native BN254 G1 multi-scalar-multiplication engine.
 *
 * Drop-in boundary for the party-local MSM of ChainSafe/co-zkvms (all paths relative to the reference tree):
 *
 *   replaces  jolt_core::msm::VariableBaseMSM::msm_field_elements(bases, gpu_bases=None, scalars, max_num_bits, use_icicle)
 *               call sites  co-jolt/src/poly/commitment/pst13.rs:286-292 (PST13::commit)
 *                           co-jolt/src/poly/commitment/pst13.rs:461-467 (open(), one MSM per variable)
 *             jolt_core::msm::VariableBaseMSM::batch_msm(bases, gpu_bases=None, polys)
 *               call site   co-jolt/src/poly/commitment/pst13.rs:319-323 (PST13::batch_commit)
 *             ark_ec::VariableBaseMSM::msm_bigint(bases, bigints)
 *               call sites  co-noir-spartan/co-spartan/src/worker.rs:804 (distributed_open)
 *                           inside MultilinearPC::commit, co-noir-spartan/co-spartan/src/worker.rs:585
 *   and the device-resident SRS the reference sketches but leaves commented out
 *               co-jolt/src/poly/commitment/pst13.rs:52-61, :235, :249 (`gpu_g1`).
 *
 * Data formats are the reference's in-memory images (arkworks 0.5):
 *   field element   4 x u64 little-endian limbs, Montgomery form (R = 2^256)            32 B
 *   base point      x || y, Fq Montgomery; `stride_bytes` lets a caller pass ark_ec's 72-byte
 *                   Affine{x, y, infinity} array without repacking; infinity flags optional
 *   scalar          Fr Montgomery (COZK_MONT: what msm_field_elements receives) or canonical
 *                   integer (COZK_CANON: the BigInt<4> msm_bigint receives); element i of a vector
 *                   lives at ptr + i*stride_bytes: 32 = dense Vec<Fr>, 64 = the `a` half of
 *                   Rep3PrimeFieldShare{a, b} (mpc-types/src/protocols/rep3/arithmetic/types.rs:22-29),
 *                   which removes copy_share_a (co-jolt/src/poly/dense_mlpoly.rs:102-110)
 *   result          72 B: x[32] || y[32] (Fq Montgomery, AFFINE, fully reduced) || infinity u8 || pad[7]
 *                   - affine because every caller normalises at once (pst13.rs:294, :328, :469; worker.rs:804),
 *                   so the bytes are independent of window size, scheduling and GPU count.
 *
 * Conventions: the caller owns every host buffer and the engine keeps no host pointer after a call returns
 * (the SRS is copied at registration).  Every function returns COZK_OK or a negative error code and never
 * unwinds; cozk_last_error() gives the message for the calling thread.  All entry points are thread-safe; calls
 * on one device serialise.  There is no CPU fallback: without a usable CUDA device cozk_init fails.
 */
#ifndef COZK_MSM_H
#define COZK_MSM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COZK_OK 0
#define COZK_ERR_INVALID_ARG (-1)    /* null pointer, k == 0, bad stride / form */
#define COZK_ERR_KEY_LENGTH (-2)     /* base_offset + n exceeds the registered SRS: the reference's
                                        ProofVerifyError::KeyLengthError / panic!("Key length error"), pst13.rs:311-316 */
#define COZK_ERR_CUDA (-3)           /* a CUDA call failed; see cozk_last_error */
#define COZK_ERR_NO_DEVICE (-4)      /* no usable CUDA device */
#define COZK_ERR_BAD_HANDLE (-5)     /* unknown SRS handle */

#define COZK_MONT 0
#define COZK_CANON 1

typedef struct cozk_ctx cozk_ctx; /* owns devices, streams, scratch */
typedef uint64_t cozk_srs;        /* handle to device-resident bases */

/* device_ids == NULL: use devices 0 .. n_devices-1 (n_devices == 0: device 0 only). */
int cozk_init(cozk_ctx** out, const int* device_ids, int n_devices);
void cozk_destroy(cozk_ctx* ctx);
int cozk_device_count(const cozk_ctx* ctx);

/* Upload n bases (replicated on every device of the context).  bases: point i at bases + i*stride_bytes,
 * stride_bytes >= 64.  infinity: n flags (non-zero = point at infinity, contributes nothing) or NULL. */
int cozk_srs_register(cozk_ctx* ctx, const void* bases, size_t n, size_t stride_bytes, const uint8_t* infinity, cozk_srs* out);
/* The same for a context with several devices when the SRS only serves point-range sharding (one long MSM, BASELINE.json
 * configs[3]): device d keeps points [n*d/D, n*(d+1)/D) only, with their own table - 1/D of the memory and of the
 * registration work on each GPU (the reference's split_ck, co-noir-spartan/co-spartan/src/utils.rs:38-83).
 * Every cozk_msm_batch over such an SRS is cut along the slices, whatever k; cozk_msm_batch_device does not take it. */
int cozk_srs_register_sliced(cozk_ctx* ctx, const void* bases, size_t n, size_t stride_bytes, const uint8_t* infinity, cozk_srs* out);
/* Drops the handle; a call that is still using the SRS on another thread keeps the device memory alive until it returns. */
int cozk_srs_release(cozk_ctx* ctx, cozk_srs srs);
int cozk_srs_len(cozk_ctx* ctx, cozk_srs srs, size_t* out_n);

/* k MSMs over the same base range [base_offset, base_offset + n) of one SRS:
 *     out[j] = sum_i scalars[j][i] * bases[base_offset + i]          j = 0 .. k-1
 * scalars: k host pointers.  max_num_bits mirrors the reference's `max_num_bits: Option<usize>`: 0 = unknown
 * (254); a smaller value promises every scalar < 2^max_num_bits and lets the engine drop high windows.
 * n == 0 gives k identities.  With several devices, k >= n_devices shards by vector, otherwise by point range
 * (the reference's split_ck + combine_comm scheme: co-noir-spartan/co-spartan/src/utils.rs:38-83,
 * snarks-core/src/poly/commitment.rs:56-63); partial sums are added on the host. */
int cozk_msm_batch(cozk_ctx* ctx, cozk_srs srs, size_t base_offset, size_t n, const void* const* scalars, size_t k,
                   size_t stride_bytes, int form, unsigned max_num_bits, void* out);

/* Same, with the k scalar vectors already in device memory of device `device_index` of the context
 * (kernel-only timing; device-resident shares kept across commit and open). */
int cozk_msm_batch_device(cozk_ctx* ctx, int device_index, cozk_srs srs, size_t base_offset, size_t n,
                          const void* const* d_scalars, size_t k, size_t stride_bytes, int form, unsigned max_num_bits,
                          void* out);

/* Ragged batch on one device: out[j] = sum_{i < lens[j]} d_scalars[j][i] * bases[base_offsets[j] + i], j = 0 .. k-1, through
 * ONE decompose / sort / accumulate / reduce - vectors of very different lengths against one SRS, e.g. the levels of a PST13
 * opening (2^(nv-1), 2^(nv-2), .. quotient scalars; pst13.rs:461-467) against the concatenated level SRSs, which as separate
 * calls pay the per-call latency nv times.  k <= 4096; returns COZK_ERR_INVALID_ARG when the group exceeds the engine's
 * bucket / pair budget (split it then). */
int cozk_msm_ragged_device(cozk_ctx* ctx, int device_index, cozk_srs srs, const size_t* base_offsets, const size_t* lens,
                           const void* const* d_scalars, size_t k, size_t stride_bytes, int form, void* out);

/* Host-side group helpers the shims need (sum of chunk commitments = combine_comm; sum of party shares =
 * PST13::combine_commitment_shares, pst13.rs:72-108).  Points are 72-byte results. */
int cozk_g1_sum(const void* points72, size_t count, void* out72);

/* Fixed-base batch multiplication, the G1 half of SRS generation:
 *     out[i] = scalars[i] * base        i = 0 .. n-1, affine
 * replaces `BatchMulPreprocessing::new(g, n).batch_mul(&pp_powers)` (co-noir-spartan/spartan/src/zk.rs:455-460) and the
 * same step inside MultilinearPC::setup behind PST13::setup (co-jolt/src/poly/commitment/pst13.rs:49-62, :276-279).
 * base72: a 72-byte wire point; scalars as in cozk_msm_batch (host pointer).  out_points72 (n x 72 B, may be NULL) gets
 * the points; out_srs (may be NULL) registers them as a device-resident SRS in the same call. */
int cozk_fixed_base_batch_mul(cozk_ctx* ctx, const void* base72, const void* scalars, size_t n, size_t stride_bytes,
                              int form, void* out_points72, cozk_srs* out_srs);

/* Tuning / introspection. */
/* "window" (0 = auto; a forced window also bypasses the SRS table), "group_pairs", "table_max_mib": memory an SRS
 * registered afterwards may spend on its precomputed table of 2^(c*w) * P rows (default 65536, and never more than half
 * of the free device memory; 0 = none).  With a table
 * all windows of a scalar share one bucket set, which removes the per-window bucket reduction and the 254 doublings of
 * the final combine; the table is built once at registration, like the reference's SRS in PST13::setup. */
/* "dominant" (default 1) / "dominant_min_points" (default 2^21): a call that covers a whole registered SRS (or an exact
 * halving of it, down to 1024 points) and has at least
 * that many (vector, point) pairs looks for windows in which (nearly) every scalar has the same digit - the constant and
 * nearly constant share vectors of co-jolt (co-jolt/src/poly/dense_mlpoly.rs:567-585) - and replaces those pairs by the
 * precomputed sum of the table row.  Same result, far fewer additions; uniform vectors pay one look at a sample.
 * "stream_min_points" (host vectors) / "chunk_min_points" (device-resident vectors; both default 2^20, 0 = never) /
 * "stream_chunks" (0 = auto): a single vector of at least that many points is cut into chunks whose H2D copies, sorts and
 * accumulate levels run as a pipeline on several streams (same result: the chunks share one bucket set).
 * "affine_rounds" (default 0 = off) / "affine_min_pairs": halving rounds of batched-affine additions (shared inversions) over
 * the sorted pair list in front of the XYZZ accumulate levels - same result; measured slower than the XYZZ kernel alone on
 * B200 (profiles/round2_affine.md), kept as an option.
 * "acc_chunk", "acc_chunk_up", "stream_first_pct", "group_l", "sort_digit_bits", "table_window", "open_small_log2", "peer_direct", "bulk_copy",
 * "chi_waves", "stream_min_points_sliced": tuning knobs, see msm.cu / engine.hpp (defaults are the measured optima). */
int cozk_set_option(cozk_ctx* ctx, const char* name, long value);
/* Timings (ms, CUDA events) of the stages of the last cozk_msm_batch* call on device 0 of the context:
 * [0] h2d  [1] decompose  [2] sort  [3] accumulate  [4] bucket-reduce  [5] finish+d2h  [6] total  [7] kernels launched
 * [8] window bits c  [9] windows W  [10] field mults (plan)  [11] pairs m */
int cozk_last_stats(cozk_ctx* ctx, double* out12);
int cozk_last_stats_device(cozk_ctx* ctx, int device_index, double* out12); /* the same for any device of the context */
const char* cozk_last_error(void);

/* ---- plain device-memory helpers so that callers (and the tests / bench) need no other CUDA binding */
int cozk_dev_alloc(cozk_ctx* ctx, int device_index, size_t bytes, void** out);
int cozk_dev_free(cozk_ctx* ctx, int device_index, void* ptr);
int cozk_dev_upload(cozk_ctx* ctx, int device_index, void* dst, const void* src, size_t bytes);
int cozk_dev_download(cozk_ctx* ctx, int device_index, void* dst, const void* src, size_t bytes);
int cozk_host_alloc_pinned(size_t bytes, void** out);
int cozk_host_free_pinned(void* ptr);
int cozk_dev_flush_l2(cozk_ctx* ctx, int device_index); /* writes a 256 MiB scratch buffer */

/* register bases that already live on device `device_index` (copied device-to-device; other devices get a peer copy) */
int cozk_srs_register_device(cozk_ctx* ctx, int device_index, const void* d_bases64, size_t n, cozk_srs* out);

#ifdef __cplusplus
}
#endif
#endif
