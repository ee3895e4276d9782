/* cozk_rep3.h - device-resident Rep3 polynomials: the steps either side of the MSM.
 *
 * A party's witness share arrives over the network, is committed (MSM), combined into one joint polynomial and opened
 * (more MSMs).  These entry points keep the share in HBM across all of that, so the only host<->device traffic of the
 * commitment path is the arrival of the share itself.  All paths below are relative to the reference tree.
 *
 *   cozk_poly_from_wire          replaces  `receive_request::<Rep3DensePolynomial>` -> deserialize_uncompressed_unchecked
 *                                          mpc-net/src/rep3/quic/worker.rs:206-219; struct co-jolt/src/poly/dense_mlpoly.rs:23-32;
 *                                          enum tag co-jolt/src/poly/multilinear_polynomial.rs:800-821
 *   cozk_poly_upload             the same for a share that already sits in host memory (in-memory Montgomery image)
 *   cozk_pst13_batch_commit_polys  PST13::batch_commit_rep3 (co-jolt/src/poly/commitment/pst13.rs:165-229) without
 *                                          copy_share_a (co-jolt/src/poly/dense_mlpoly.rs:102-110) and without H2D
 *   cozk_rep3_linear_combination Rep3MultilinearPolynomial::linear_combination
 *                                          co-jolt/src/poly/multilinear_polynomial.rs:196-296 (call site opening_proof.rs:274-278)
 *   cozk_rep3_evaluate_at_chi    Rep3DensePolynomial::evaluate_at_chi / batch_evaluate  co-jolt/src/poly/dense_mlpoly.rs:160-194
 *   cozk_eq_evals                the eq table those dot products need, built on the device (jolt-core EqPolynomial::evals
 *                                          order, or ark-poly's evaluate / fix_variables order)
 *   cozk_spartan_batch_open_worker  distributed_batch_open_poly_worker  co-noir-spartan/co-spartan/src/worker.rs:745-772
 *                                          (aggregate_poly + distributed_open + per-polynomial evaluate)
 *   cozk_pst13_open_key_create + cozk_pst13_open_poly   open() behind PST13::prove_rep3 (pst13.rs:125-137, :428-474) on the
 *                                          device-resident joint polynomial; every quotient scalar multiplies two adjacent
 *                                          bases (pst13.rs:459), so level i runs as a half-size MSM over P[2b] + P[2b+1]
 *                                          (cozk_srs_pair_sums), and the many small levels run as one batched MSM.
 *
 * Formats: Fr in memory = 4 x u64 LE limbs, Montgomery (arkworks); Fr on the wire = 32-byte LE canonical integer
 * (ark-serialize, uncompressed).  A shared polynomial is an array of Rep3PrimeFieldShare{a, b}
 * (mpc-types/src/protocols/rep3/arithmetic/types.rs:22-29): 64 B per coefficient.
 */
#ifndef COZK_REP3_H
#define COZK_REP3_H
#include "cozk_msm.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t cozk_poly; /* handle to a device-resident polynomial */

/* kinds (host image passed to cozk_poly_upload) */
#define COZK_POLY_SHARED 0 /* Rep3DensePolynomial: len x {a, b}, Fr Montgomery, 64 B each */
#define COZK_POLY_PUBLIC 1 /* MultilinearPolynomial::LargeScalars: len x Fr Montgomery, 32 B each */
#define COZK_POLY_U8 2     /* MultilinearPolynomial::U8Scalars ... I64Scalars (multilinear_polynomial.rs:226-268): */
#define COZK_POLY_U16 3    /*   len x 1 / 2 / 4 / 8 bytes, little-endian; widened on the device */
#define COZK_POLY_U32 4
#define COZK_POLY_U64 5
#define COZK_POLY_I64 6

#define COZK_ERR_WIRE (-6) /* malformed or truncated wire image, or a field element >= r (ark: SerializationError) */

/* Upload the in-memory image of a polynomial to device `device_index` of the context. */
int cozk_poly_upload(cozk_ctx* ctx, int device_index, const void* coeffs, size_t len, int kind, cozk_poly* out);

/* The same from an image that already sits on that device (SHARED or PUBLIC only); copied device to device. */
int cozk_poly_from_device(cozk_ctx* ctx, int device_index, const void* d_coeffs, size_t len, int kind, cozk_poly* out);

/* Parse one ark-serialize (uncompressed) Rep3DensePolynomial from `bytes` and keep it on the device: the coefficient block
 * goes to HBM as it is and is converted canonical -> Montgomery there.  tagged != 0: the image starts with the
 * Rep3MultilinearPolynomial discriminant byte, which must be 1 (Shared); the Public variant embeds jolt-core's
 * MultilinearPolynomial, whose wire format is not defined in the reference tree.  chunk_range is honoured (the handle
 * covers coeffs[chunk_range.0 .. chunk_range.1], as copy_share_a does); bound_coeffs / binding_scratch_space are skipped.
 * consumed (may be NULL) gets the number of bytes read. */
int cozk_poly_from_wire(cozk_ctx* ctx, int device_index, const void* bytes, size_t nbytes, int tagged, cozk_poly* out,
                        size_t* consumed);

int cozk_poly_release(cozk_ctx* ctx, cozk_poly poly);
/* len: coefficients; kind: COZK_POLY_SHARED, COZK_POLY_PUBLIC, or the small-scalar kind it was uploaded as */
int cozk_poly_info(cozk_ctx* ctx, cozk_poly poly, size_t* len, int* kind, int* device_index);
/* Device image back to the host: SHARED len x 64 B, PUBLIC len x 32 B (both Montgomery), small kinds len x 32 B canonical. */
int cozk_poly_download(cozk_ctx* ctx, cozk_poly poly, void* out);

/* PST13::batch_commit_rep3 over device-resident polynomials of one length n = 2^nv.  Shared polynomials commit share a;
 * public ones are committed only when commit_to_public != 0 (party 0).  Outputs as cozk_pst13_batch_commit_rep3. */
int cozk_pst13_batch_commit_polys(cozk_ctx* ctx, cozk_srs srs, const cozk_poly* polys, size_t k, int commit_to_public,
                                  void* out_commitments, uint8_t* present);

/* joint[i] = sum_{j : i < len_j} coeffs[j] * polys[j][i],  i < max_j len_j.  coeffs: k Fr values, Montgomery.
 * With at least one shared input the result is shared: public terms join share a on party 0, share b on party 1 and
 * neither on party 2 (rep3 add_public); every index must then be covered by a shared polynomial (the reference panics in
 * `as_shared()` otherwise) -> COZK_ERR_INVALID_ARG.  With public inputs only the result is a public polynomial.
 * Inputs on one device: the result lives there too.  Inputs spread over several devices of the context (a party's
 * polynomials dealt over the GPUs of a box, each device committing its own): every device forms the combination of ITS
 * polynomials, and one kernel on device 0 - where the opening runs - adds the partial results, reading the remote ones
 * through NVLink peer mappings (staged peer copies when two devices have no peer path, or with option "peer_direct" = 0);
 * the result lives on device 0. */
int cozk_rep3_linear_combination(cozk_ctx* ctx, const cozk_poly* polys, const void* coeffs, size_t k, int party_id,
                                 cozk_poly* out);

/* out[j] = sum_i into_additive(polys[j][i]) * chis[i] = TWO_INV * sum_i (a_i + b_i) * chis[i] for a shared polynomial
 * (an AdditiveShare), sum_i v_i * chis[i] for a public one.  chis: n Fr values, Montgomery, host memory; n must equal
 * every polynomial's length (zip_eq).  out: k x 32 B, Montgomery. */
int cozk_rep3_evaluate_at_chi(cozk_ctx* ctx, const cozk_poly* polys, size_t k, const void* chis, size_t n, void* out_evals);

/* The same with the chi table already on the device (a public polynomial handle, e.g. from cozk_eq_evals). */
int cozk_rep3_evaluate_at_chi_poly(cozk_ctx* ctx, const cozk_poly* polys, size_t k, cozk_poly chi, void* out_evals);

/* eq table of a point as a device-resident public polynomial of 2^nv values:
 *   chi[b] = prod_i (bit_i(b) ? t_i : 1 - t_i),  t_i = point[i] (msb_first == 0) or point[nv-1-i] (msb_first != 0).
 * msb_first == 0 is the order in which ark-poly's DenseMultilinearExtension::evaluate / fix_variables and open()
 * (pst13.rs:454-458, worker.rs:793-798) consume the point: sum_b poly[b] * chi[b] == poly.evaluate(point).
 * msb_first != 0 is jolt-core's EqPolynomial::evals(r), the `chis` of Rep3DensePolynomial::evaluate_at_chi
 * (co-jolt/src/poly/dense_mlpoly.rs:160-181).  point: nv Fr values, Montgomery, host memory. */
int cozk_eq_evals(cozk_ctx* ctx, int device_index, const void* point, size_t nv, int msb_first, cozk_poly* out);

/* out = SRS of n/2 points S[b] = P[2b] + P[2b+1] (n even).  Setup-time work, like the SRS itself. */
int cozk_srs_pair_sums(cozk_ctx* ctx, cozk_srs srs, cozk_srs* out);

/* Setup-time key of the PST13 opening: the pair sums of every level (each MSM halves), with the levels that hold at most
 * 2^15 quotient scalars (option "open_small_log2") concatenated into one SRS so that ONE batched MSM opens all of them -
 * an MSM of a few thousand points is latency-bound, a dozen of them in a row dominate the reference's schedule.
 * level_srs[i] = ck.powers_of_g[i] (2^(nv-i) points) stay owned by the caller and must outlive the key. */
typedef uint64_t cozk_open_key;
int cozk_pst13_open_key_create(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv, cozk_open_key* out);
int cozk_pst13_open_key_release(cozk_ctx* ctx, cozk_open_key key);

/* PST13 opening of a device-resident shared polynomial (share a) or public polynomial of 2^nv coefficients on device 0.
 * key != 0: the keyed schedule above (level_srs / nv are taken from the key); key == 0: the reference's schedule, one MSM
 * per level over the duplicated quotient scalars.  All folds run on the device; nothing but the nv proof points and the
 * evaluation comes back.  point / out_proofs / out_eval as cozk_pst13_open (cozk_pst13.h). */
int cozk_pst13_open_poly(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv, cozk_open_key key, cozk_poly poly,
                         const void* point, void* out_proofs, void* out_eval);

/* cozk_pst13_open with a key: host evaluations as in cozk_pst13_open. */
int cozk_pst13_open_keyed(cozk_ctx* ctx, cozk_open_key key, const void* evals, size_t stride_bytes, const void* point,
                          int form, void* out_proofs, void* out_eval);

/* co-spartan's distributed_batch_open_poly_worker (co-noir-spartan/co-spartan/src/worker.rs:745-772) without the network
 * send, on device 0 from device-resident polynomials (the party's share_0 evaluations, COZK_POLY_PUBLIC images):
 *   agg = aggregate_poly(eta, polys[0..num_comms]) (co-spartan/src/utils.rs:85-107); (proofs, val) = distributed_open(ck,
 *   agg, point) (worker.rs:774-809); evals[j] = polys[j].evaluate(point) for ALL k polynomials (:761-764).
 * key != 0: the keyed opening schedule (level_srs / nv taken from the key); else level_srs[i] = ck.powers_of_g[i].
 * point: nv Fr, eta: one Fr (Montgomery).  out_proofs: nv x 72 B; out_val: 32 B; out_evals: k x 32 B (Montgomery).
 * Every polynomial must have 2^nv evaluations (evaluate asserts the point length; COZK_ERR_INVALID_ARG / KEY_LENGTH). */
int cozk_spartan_batch_open_worker(cozk_ctx* ctx, cozk_open_key key, const cozk_srs* level_srs, size_t nv, const cozk_poly* polys,
                                   size_t k, size_t num_comms, const void* point, const void* eta, void* out_proofs,
                                   void* out_val, void* out_evals);

/* Timings (ms, CUDA events on the engine's stream) of the last call on this thread's context:
 * [0] cozk_poly_from_wire / upload: H2D  [1] ingest kernel  [2] linear combination kernel  [3] chi kernels  [4] bytes moved by [2] */
int cozk_rep3_last_stats(cozk_ctx* ctx, double* out8);

#ifdef __cplusplus
}
#endif
#endif
