/* cozk_test.h - TEST AND MEASUREMENT entry points of libcozk_msm.so: synthetic inputs, element-wise kernels for parity
 * tests, the pair sort on its own, roofline microbenchmarks.  Nothing here is part of the drop-in boundary
 * (include/cozk_msm.h, cozk_rep3.h, cozk_pst13.h); tests/, bench.py and tools/ are the only users. */
#ifndef COZK_TEST_H
#define COZK_TEST_H
#include "cozk_msm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic inputs, generated on the device (bit-identical to oracle/bn254.c) */
int cozk_testgen_bases(cozk_ctx* ctx, int device_index, uint64_t seed, size_t start, size_t n, void* d_out64);
int cozk_testgen_scalars(cozk_ctx* ctx, int device_index, int dist, uint64_t seed, size_t start, size_t n, size_t total_n,
                         int form, void* d_out, size_t stride_bytes);
/* ---- element-wise kernels exposed for parity tests and roofline microbenchmarks (device pointers) */
/* op: 0 fq_mul 1 fq_add 2 fq_sub 3 fq_sqr 4 fq_inv 5 fr_from_mont; arrays of n 32-byte elements */
int cozk_test_field_op(cozk_ctx* ctx, int device_index, int op, const void* d_a, const void* d_b, void* d_out, size_t n);
/* op: 0 xyzz_add 1 xyzz_madd 2 xyzz_dbl on arrays of n 72-byte wire points */
int cozk_test_g1_op(cozk_ctx* ctx, int device_index, int op, const void* d_a, const void* d_b, void* d_out, size_t n);
/* which: 0 = independent IMAD.WIDE chains (pipe peak), 1 = dependent fq_mul chains, 2 = fq_sqr chains, 3 = xyzz_madd chain,
 * 4 = IMAD.WIDE carry chains (mad.lo.cc/madc.hi.cc rows), 5 = mad.lo.u32, 6 = mad.hi.u32, 7 = four fq_mul chains per thread.
 * Runs `iters` operations per thread on blocks x threads; returns elapsed ms and the operation count. */
int cozk_microbench(cozk_ctx* ctx, int device_index, int which, int blocks, int threads, int iters, double* out_ms,
                    double* out_ops);

/* The pair sort on its own (csrc/sort_kernels.cuh).  d_scalars == NULL: groups the m given (key, val) pairs by key, keys
 * ascending (every key below 2^key_bits; the order inside a group is unspecified).  d_scalars != NULL: the pairs are those of the plain decompose layout of g vectors
 * of n scalars (vector v at d_scalars + v * round_up((n-1)*stride + 32, 256)); fused != 0 produces them inside the first
 * sort pass (the engine's path), fused == 0 with the decompose kernel followed by generic passes (key_bits == 0: left
 * unsorted).  Outputs hold m = g * n * windows pairs. */
int cozk_test_sort(cozk_ctx* ctx, int device_index, const void* d_keys, const void* d_vals, size_t m, unsigned key_bits,
                   const void* d_scalars, size_t n, unsigned g, size_t stride, int form, unsigned c, unsigned windows,
                   size_t table_stride, size_t val_offset, int fused, void* d_keys_out, void* d_vals_out);

/* The batched-affine pre-reduction on its own (csrc/affine_kernels.cuh): `rounds` halving rounds over m pairs grouped by key
 * (vals index into d_pts: 64-byte affine points; bit 31 = negate; all ones = skip).  reference == 0: the engine's batched kernels,
 * != 0: the serial contract kernel.  Outputs (device pointers, may be NULL): the reduced list of out_m[0] entries, and per round r the
 * overflow list at d_ovf_keys + r * ovf_stride entries / d_ovf_pts + r * ovf_stride points, ovf_counts[r] entries (host array).
 * out_ms: time of the rounds. */
int cozk_test_affine_rounds(cozk_ctx* ctx, int device_index, const void* d_keys, const void* d_vals, size_t m, const void* d_pts,
                            size_t total_buckets, int rounds, int reference, void* d_keys_out, void* d_vals_out, void* d_pts_out,
                            size_t* out_m, void* d_ovf_keys, void* d_ovf_pts, size_t ovf_stride, unsigned* ovf_counts, double* out_ms);

#ifdef __cplusplus
}
#endif
#endif
