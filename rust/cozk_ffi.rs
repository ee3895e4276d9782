//! UN-COMPILED REFERENCE MATERIAL (no Rust toolchain exists in the build environment).
//!
//! Drop this file into `co-jolt/src/poly/commitment/cozk.rs` (and `co-noir-spartan/co-spartan/src/cozk.rs`), add
//! `mod cozk;`, link `libcozk_msm.so` from a `build.rs`, and swap the call sites listed in INTEGRATION.md section 2.
//! Every `extern "C"` item below is declared in `include/cozk_msm.h` / `include/cozk_pst13.h` and exported by the
//! library (checked by `tests/test_abi.py`).

#![allow(dead_code)]
use std::os::raw::{c_char, c_int, c_long, c_uint, c_void};

#[repr(C)]
pub struct CozkCtx {
    _private: [u8; 0],
}
pub type CozkSrs = u64;

pub const COZK_OK: c_int = 0;
pub const COZK_ERR_INVALID_ARG: c_int = -1;
pub const COZK_ERR_KEY_LENGTH: c_int = -2;
pub const COZK_ERR_CUDA: c_int = -3;
pub const COZK_ERR_NO_DEVICE: c_int = -4;
pub const COZK_ERR_BAD_HANDLE: c_int = -5;
pub const COZK_POLY_SHARED: c_int = 0; // include/cozk_rep3.h
pub const COZK_POLY_PUBLIC: c_int = 1;
pub const COZK_POLY_U8: c_int = 2;
pub const COZK_POLY_U16: c_int = 3;
pub const COZK_POLY_U32: c_int = 4;
pub const COZK_POLY_U64: c_int = 5;
pub const COZK_POLY_I64: c_int = 6;
pub const COZK_MONT: c_int = 0;
pub const COZK_CANON: c_int = 1;

#[link(name = "cozk_msm")]
extern "C" {
    pub fn cozk_init(out: *mut *mut CozkCtx, device_ids: *const c_int, n_devices: c_int) -> c_int;
    pub fn cozk_destroy(ctx: *mut CozkCtx);
    pub fn cozk_device_count(ctx: *const CozkCtx) -> c_int;
    pub fn cozk_srs_register(ctx: *mut CozkCtx, bases: *const c_void, n: usize, stride_bytes: usize, infinity: *const u8, out: *mut CozkSrs) -> c_int;
    /// every device keeps only its point range (+ its table): for an SRS that only serves one long MSM sharded by point range
    pub fn cozk_srs_register_sliced(ctx: *mut CozkCtx, bases: *const c_void, n: usize, stride_bytes: usize, infinity: *const u8, out: *mut CozkSrs) -> c_int;
    pub fn cozk_srs_release(ctx: *mut CozkCtx, srs: CozkSrs) -> c_int;
    pub fn cozk_srs_len(ctx: *mut CozkCtx, srs: CozkSrs, out_n: *mut usize) -> c_int;
    pub fn cozk_msm_batch(ctx: *mut CozkCtx, srs: CozkSrs, base_offset: usize, n: usize, scalars: *const *const c_void, k: usize,
                          stride_bytes: usize, form: c_int, max_num_bits: c_uint, out: *mut c_void) -> c_int;
    pub fn cozk_g1_sum(points72: *const c_void, count: usize, out72: *mut c_void) -> c_int;
    pub fn cozk_fixed_base_batch_mul(ctx: *mut CozkCtx, base72: *const c_void, scalars: *const c_void, n: usize, stride_bytes: usize,
                                     form: c_int, out_points72: *mut c_void, out_srs: *mut CozkSrs) -> c_int;
    pub fn cozk_set_option(ctx: *mut CozkCtx, name: *const c_char, value: c_long) -> c_int;
    pub fn cozk_last_error() -> *const c_char;
    // include/cozk_pst13.h
    pub fn cozk_pst13_commit(ctx: *mut CozkCtx, srs: CozkSrs, evals: *const c_void, n: usize, stride_bytes: usize, form: c_int,
                             max_num_bits: c_uint, out_commitment: *mut c_void) -> c_int;
    pub fn cozk_pst13_batch_commit(ctx: *mut CozkCtx, srs: CozkSrs, polys: *const *const c_void, k: usize, n: usize, stride_bytes: usize,
                                   form: c_int, max_num_bits: *const c_uint, out_commitments: *mut c_void) -> c_int;
    pub fn cozk_pst13_batch_commit_rep3(ctx: *mut CozkCtx, srs: CozkSrs, polys: *const *const c_void, is_shared: *const u8, k: usize,
                                        n: usize, form: c_int, max_num_bits: *const c_uint, commit_to_public: c_int,
                                        out_commitments: *mut c_void, present: *mut u8) -> c_int;
    /// polys[j] in its in-memory form: kinds[j] = COZK_POLY_SHARED (Rep3PrimeFieldShare array), COZK_POLY_PUBLIC (LargeScalars) or
    /// COZK_POLY_U8 .. COZK_POLY_I64 (MultilinearPolynomial::U8Scalars .. I64Scalars as they lie in memory; widened on the device)
    pub fn cozk_pst13_batch_commit_packed(ctx: *mut CozkCtx, srs: CozkSrs, polys: *const *const c_void, kinds: *const c_int, k: usize,
                                          n: usize, commit_to_public: c_int, out_commitments: *mut c_void, present: *mut u8) -> c_int;
    pub fn cozk_pst13_open(ctx: *mut CozkCtx, level_srs: *const CozkSrs, nv: usize, evals: *const c_void, stride_bytes: usize,
                           point: *const c_void, form: c_int, out_proofs: *mut c_void, out_eval: *mut c_void) -> c_int;
    pub fn cozk_pst13_combine_commitment_shares(commitments: *const c_void, count: usize, out_commitment: *mut c_void) -> c_int;
    pub fn cozk_pst13_coordinate_prove(proofs: *const c_void, parties: usize, len: usize, out_proofs: *mut c_void) -> c_int;
    pub fn cozk_combine_comm(commitments: *const c_void, count: usize, out_commitment: *mut c_void) -> c_int;
}

/// 72-byte result of the engine: x || y are ark_ff's in-memory Montgomery limbs.
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct WirePoint {
    pub x: [u64; 4],
    pub y: [u64; 4],
    pub infinity: u8,
    pub _pad: [u8; 7],
}

/// PST13Commitment{nv, g_product} as the C++ layer lays it out (80 bytes).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct WireCommitment {
    pub nv: u64,
    pub g_product: WirePoint,
}

impl WirePoint {
    pub fn into_affine(self) -> ark_bn254::G1Affine {
        use ark_ec::AffineRepr;
        if self.infinity != 0 {
            return ark_bn254::G1Affine::identity();
        }
        let x = ark_bn254::Fq::new_unchecked(ark_ff::BigInt(self.x));
        let y = ark_bn254::Fq::new_unchecked(ark_ff::BigInt(self.y));
        ark_bn254::G1Affine::new_unchecked(x, y)
    }
    pub fn from_affine(p: &ark_bn254::G1Affine) -> Self {
        use ark_ec::AffineRepr;
        match p.xy() {
            None => WirePoint { infinity: 1, ..Default::default() },
            Some((x, y)) => WirePoint { x: (x.0).0, y: (y.0).0, infinity: 0, _pad: [0; 7] },
        }
    }
}

/// Replacement for `jolt_core::msm::VariableBaseMSM::msm_field_elements(bases, None, scalars, max_num_bits, _)`
/// (call sites co-jolt/src/poly/commitment/pst13.rs:286, :461): same inputs, same `Result`, the SRS level is already
/// registered on the device.
pub fn msm_field_elements(ctx: *mut CozkCtx, srs: CozkSrs, len: usize, scalars: &[ark_bn254::Fr], max_num_bits: Option<usize>)
    -> Result<ark_bn254::G1Projective, jolt_core::utils::errors::ProofVerifyError> {
    use jolt_core::utils::errors::ProofVerifyError;
    if scalars.len() != len {
        return Err(ProofVerifyError::KeyLengthError(len, scalars.len()));
    }
    let mut out = WirePoint::default();
    let ptr = [scalars.as_ptr() as *const c_void];
    let rc = unsafe {
        cozk_msm_batch(ctx, srs, 0, len, ptr.as_ptr(), 1, 32, COZK_MONT, max_num_bits.unwrap_or(0) as c_uint,
                       &mut out as *mut _ as *mut c_void)
    };
    match rc {
        COZK_OK => Ok(out.into_affine().into()),
        COZK_ERR_KEY_LENGTH => Err(ProofVerifyError::KeyLengthError(len, scalars.len())),
        _ => panic!("cozk_msm_batch failed: {}", unsafe { std::ffi::CStr::from_ptr(cozk_last_error()) }.to_string_lossy()),
    }
}

/// Replacement for `batch_msm(bases, None, polys)` on the shared polynomials of `batch_commit_rep3`
/// (pst13.rs:165-229): share `a` is read in place at stride 64, `copy_share_a` disappears.
pub fn batch_msm_rep3_share_a(ctx: *mut CozkCtx, srs: CozkSrs, len: usize,
                              shares: &[&[mpc_core::protocols::rep3::Rep3PrimeFieldShare<ark_bn254::Fr>]]) -> Vec<ark_bn254::G1Affine> {
    let ptrs: Vec<*const c_void> = shares.iter().map(|s| &s[0].a as *const _ as *const c_void).collect();
    let mut out = vec![WirePoint::default(); shares.len()];
    let rc = unsafe {
        cozk_msm_batch(ctx, srs, 0, len, ptrs.as_ptr(), ptrs.len(), 64, COZK_MONT, 0, out.as_mut_ptr() as *mut c_void)
    };
    if rc == COZK_ERR_KEY_LENGTH {
        panic!("Key length error"); // pst13.rs:311-316
    }
    assert_eq!(rc, COZK_OK);
    out.into_iter().map(WirePoint::into_affine).collect()
}

/// Replacement for `ark_ec::VariableBaseMSM::msm_bigint(bases, bigints)` (co-noir-spartan/co-spartan/src/worker.rs:804).
pub fn msm_bigint(ctx: *mut CozkCtx, srs: CozkSrs, bigints: &[ark_ff::BigInt<4>]) -> ark_bn254::G1Projective {
    let mut out = WirePoint::default();
    let ptr = [bigints.as_ptr() as *const c_void];
    let mut len = 0usize;
    unsafe { cozk_srs_len(ctx, srs, &mut len) };
    let n = bigints.len().min(len); // arkworks truncates to the shorter slice
    let rc = unsafe { cozk_msm_batch(ctx, srs, 0, n, ptr.as_ptr(), 1, 32, COZK_CANON, 0, &mut out as *mut _ as *mut c_void) };
    assert_eq!(rc, COZK_OK);
    out.into_affine().into()
}
