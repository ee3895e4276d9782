/* cozk_pst13.h - host-side mirror of the reference's commitment-scheme interface for the MSM path, above the
 * MSM C ABI (include/cozk_msm.h).  Same operations, argument meaning and error behaviour as
 *
 *   PST13::commit                      co-jolt/src/poly/commitment/pst13.rs:282-296
 *   PST13::batch_commit                co-jolt/src/poly/commitment/pst13.rs:299-331
 *   PST13::batch_commit_rep3           co-jolt/src/poly/commitment/pst13.rs:165-229   (commit_rep3: :140-162)
 *   open() behind PST13::prove_rep3    co-jolt/src/poly/commitment/pst13.rs:428-474   (:125-137)
 *   PST13::combine_commitment_shares   co-jolt/src/poly/commitment/pst13.rs:72-108
 *   poly_commit_worker / distributed_open (co-spartan)  co-noir-spartan/co-spartan/src/worker.rs:577-590, :774-809
 *     - the same two operations over `z.share_0.evaluations` (dense Vec<Fr>), so they map onto
 *       cozk_pst13_commit / cozk_pst13_open with stride 32.
 *
 * The Rust toolchain is absent in this environment, so this layer is C++ behind extern "C" (the reference is
 * compiled code); INTEGRATION.md shows the Rust shim that binds it.
 *
 * A commitment is the reference's PST13Commitment{nv, g_product} (pst13.rs:398-401) laid out as
 *   u64 nv || 72-byte wire point                                                        (80 B)
 */
#ifndef COZK_PST13_H
#define COZK_PST13_H
#include "cozk_msm.h"

#ifdef __cplusplus
extern "C" {
#endif

#define COZK_COMMITMENT_BYTES 80

/* g_product = sum_i evals[i] * srs[i], i < n; nv = log2(n) (n must be a power of two, as get_num_vars assumes). */
int cozk_pst13_commit(cozk_ctx* ctx, cozk_srs srs, const void* evals, size_t n, size_t stride_bytes, int form,
                      unsigned max_num_bits, void* out_commitment);

/* k polynomials of one length n against srs[0..n).  max_num_bits: k hints (MultilinearPolynomial::{U8,U16,U32,U64}Scalars
 * -> 8/16/32/64, LargeScalars / I64Scalars -> 0) or NULL.  n > srs length -> COZK_ERR_KEY_LENGTH ("Key length error"). */
int cozk_pst13_batch_commit(cozk_ctx* ctx, cozk_srs srs, const void* const* polys, size_t k, size_t n, size_t stride_bytes,
                            int form, const unsigned* max_num_bits, void* out_commitments);

/* Rep3 variant.  is_shared[j] != 0: polys[j] points at share `a` of element 0 of an AoS Rep3PrimeFieldShare{a, b}
 * array (stride 64; copy_share_a is not needed); is_shared[j] == 0: a public polynomial, dense stride 32, committed
 * only when commit_to_public != 0 (party 0) - the reference computes those MSMs on every party and drops the result on
 * parties 1/2 (pst13.rs:198-225); skipping them changes no output.  present[j] = 1 for MaybeShared::Shared(_) and
 * MaybeShared::Public(Some(_)), 0 for MaybeShared::Public(None). */
int cozk_pst13_batch_commit_rep3(cozk_ctx* ctx, cozk_srs srs, const void* const* polys, const uint8_t* is_shared, size_t k,
                                 size_t n, int form, const unsigned* max_num_bits, int commit_to_public,
                                 void* out_commitments, uint8_t* present);

/* The same for polynomials in the reference's PACKED in-memory forms: kinds[j] is one of COZK_POLY_* (cozk_rep3.h) -
 * COZK_POLY_SHARED (n x Rep3PrimeFieldShare{a, b}, 64 B), COZK_POLY_PUBLIC (MultilinearPolynomial::LargeScalars, n x Fr
 * Montgomery) or COZK_POLY_U8 / U16 / U32 / U64 / I64 (MultilinearPolynomial::U8Scalars .. I64Scalars,
 * co-jolt/src/poly/multilinear_polynomial.rs:226-268: n x 1 / 2 / 4 / 8 bytes).  The small-scalar polynomials - about 60 of
 * the ~198 trace polynomials - cross PCIe as they lie in memory (1 - 8 bytes per coefficient instead of 32) and are widened
 * on the device; `batch_msm` dispatches on the same enum (pst13.rs:319-323).  Public kinds are committed only when
 * commit_to_public != 0; present[] as above.  With several devices the polynomials are dealt round-robin and the devices
 * commit side by side (the SRS must be a replicated one). */
int cozk_pst13_batch_commit_packed(cozk_ctx* ctx, cozk_srs srs, const void* const* polys, const int* kinds, size_t k, size_t n,
                                   int commit_to_public, void* out_commitments, uint8_t* present);

/* PST13 opening: level_srs[i] holds ck.powers_of_g[i] (2^(nv-i) points), i < nv.  evals: 2^nv Fr values (Montgomery)
 * at stride_bytes; point: nv Fr values (Montgomery) in the order open() receives them (prove_rep3 reverses the
 * opening point first, pst13.rs:134).  out_proofs: nv wire points; out_eval: r[0][0] (Fr Montgomery, 32 B). */
int cozk_pst13_open(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv, const void* evals, size_t stride_bytes,
                    const void* point, int form, void* out_proofs, void* out_eval);

/* Coordinator side: sum of the parties' commitment shares; asserts equal nv (COZK_ERR_INVALID_ARG otherwise). */
int cozk_pst13_combine_commitment_shares(const void* commitments, size_t count, void* out_commitment);

/* Coordinator side of prove_rep3: PST13::coordinate_prove (pst13.rs:110-122) - element-wise a + b + c of the parties'
 * proof vectors.  proofs: `parties` arrays of `len` wire points each, party-major; out: len wire points. */
int cozk_pst13_coordinate_prove(const void* proofs, size_t parties, size_t len, void* out_proofs);

/* combine_comm (snarks-core/src/poly/commitment.rs:56-63): sum of the workers' chunk commitments,
 * nv = nv_0 + log2(count); count must be a power of two as `log_2()` assumes. */
int cozk_combine_comm(const void* commitments, size_t count, void* out_commitment);

#ifdef __cplusplus
}
#endif
#endif
