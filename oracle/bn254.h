/* oracle/bn254.h - CPU oracle for the party-local BN254 G1 MSM path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (co-zkvms_b200/) may include, link or call
 * this.  Users: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and --impl reference).
 *
 * PARITY UNPINNED.  The reference's MSM lives in un-vendored Rust crates (jolt-core 0.1.0 @
 * nulltea/jolt cd50b476; ark-ec / ark-ff 0.5.0 @ a16z/arkworks-algebra 4ae5018; ark-bn254 0.5.0) and
 * there is no Rust toolchain here, so this is a restatement of the published algorithm, anchored on the
 * reference's call sites (co-jolt/src/poly/commitment/pst13.rs:282-331, :428-474;
 * co-noir-spartan/co-spartan/src/worker.rs:577-590, :774-809) and checked against
 *  (1) oracle/pyref.py (Python big-int affine arithmetic) through tests/golden/ fixtures,
 *  (2) the public EIP-196 known answers for BN254 add / mul,
 *  (3) algebraic identities that hold for any correct MSM.
 *
 * Wire formats are those of include/cozk_msm.h:
 *   field element  = 4 x u64 little-endian limbs (arkworks BigInt<4> in-memory layout)
 *   base point     = x[32] || y[32], Fq Montgomery form (64 B)
 *   scalar         = 32 B, Fr Montgomery (ORC_MONT) or canonical integer (ORC_CANON)
 *   result         = x[32] || y[32] (Fq Montgomery, affine, normalised) || infinity u8 || pad[7]  (72 B)
 */
#ifndef ORACLE_BN254_H
#define ORACLE_BN254_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MONT 0
#define ORC_CANON 1

/* scalar distributions; must match oracle/pyref.py:scalars() */
enum { ORC_DIST_UNIFORM = 0, ORC_DIST_CONST = 1, ORC_DIST_WMINUS = 2, ORC_DIST_DUP = 3, ORC_DIST_SMALL16 = 4, ORC_DIST_ZERO_HALF = 5 };

/* ---- element-wise field ops on arrays of n elements (32 B each), Montgomery form in and out.
 * which: 0 = Fq (base field), 1 = Fr (scalar field).  op: 0 add, 1 sub, 2 mul, 3 sqr(a), 4 neg(a), 5 inv(a) (0 -> 0),
 * 6 to_mont(a), 7 from_mont(a). */
void orc_field_op(int which, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);

/* ---- curve ops on arrays of n affine points in the 72-byte result format.
 * op: 0 add, 1 double(a), 2 neg(a).  Goes through the Jacobian formulas and normalises. */
void orc_g1_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
/* returns 1 iff the 72-byte point is the identity or satisfies y^2 = x^3 + 3 with canonical coordinates */
int orc_g1_is_valid(const uint8_t* pt72);
/* scalar * point, double-and-add on Jacobian coordinates */
void orc_g1_mul(const uint8_t* pt72, const uint64_t* scalar, int form, uint8_t* out72);

/* ---- deterministic synthetic inputs (identical to oracle/pyref.py and csrc/testgen.cu) */
void orc_gen_bases(uint64_t seed, size_t start, size_t n, uint8_t* out64, int threads);
void orc_gen_scalars(int dist, uint64_t seed, size_t start, size_t n, size_t total_n, int form, uint8_t* out, size_t stride_bytes);

/* ---- the MSM itself.
 * orc_msm_naive: sum of per-point double-and-add; O(254 n) group ops; the slow cross-check.
 * orc_msm: restatement of ark-ec 0.5 VariableBaseMSM::msm_bigint_wnaf (window c = 3 if n < 32 else
 *   ceil(log2 n)*69/100 + 2; signed base-2^c digits with the last digit absorbing the final carry;
 *   per-window bucket fill with mixed additions; running-sum bucket reduction; high-to-low window
 *   combine with c doublings each).  threads <= 1 runs exactly that, windows in sequence.  threads > 1
 *   parallelises over windows (as arkworks does with rayon) and additionally over point chunks so that
 *   more cores than windows are busy (as jolt-core's batch_msm does across polynomials).
 * scalars: element i at scalars + i*stride_bytes (32 = dense Vec<Fr>, 64 = Rep3 AoS share a). */
void orc_msm_naive(const uint8_t* bases64, const uint8_t* scalars, size_t stride_bytes, int form, size_t n, uint8_t* out72);
void orc_msm(const uint8_t* bases64, const uint8_t* scalars, size_t stride_bytes, int form, size_t n, int threads, uint8_t* out72);
/* window size arkworks would pick for n points */
int orc_msm_window(size_t n);
/* number of group operations (mixed adds, full adds, doublings) of the last orc_msm call on this thread */
void orc_msm_last_counts(uint64_t* madd, uint64_t* add, uint64_t* dbl);

#ifdef __cplusplus
}
#endif
#endif
