"""Python face of include/cozk_pst13.h: the reference's PST13 commitment-scheme operations that reach the MSM.

Names follow co-jolt/src/poly/commitment/pst13.rs so that the parity tests read like the reference's own test
(pst13.rs:476-547): PST13Setup, commit, batch_commit, batch_commit_rep3, open, combine_commitment_shares.
Field elements are 32-byte little-endian Montgomery values (arkworks' in-memory Fr / Fq); a commitment is
(nv, 72-byte wire point).
"""
import ctypes

import numpy as np

COMMITMENT_BYTES = 80


def _lib():
    from . import lib
    return lib()


def _check(rc):
    from . import _check as chk
    chk(rc)


class PST13Commitment:
    """PST13Commitment{nv, g_product} (pst13.rs:398-401)."""

    def __init__(self, nv, g_product):
        self.nv = int(nv)
        self.g_product = np.array(g_product, dtype=np.uint8).reshape(72)

    @classmethod
    def from_bytes(cls, raw):
        raw = np.asarray(raw, dtype=np.uint8)
        return cls(int.from_bytes(bytes(raw[:8]), "little"), raw[8:80])

    def to_bytes(self):
        out = np.zeros(COMMITMENT_BYTES, dtype=np.uint8)
        out[:8] = np.frombuffer(int(self.nv).to_bytes(8, "little"), dtype=np.uint8)
        out[8:] = self.g_product
        return out

    def __eq__(self, other):
        return self.nv == other.nv and bool((self.g_product == other.g_product).all())


class PST13Setup:
    """Device-resident `uni_params.powers_of_g` (pst13.rs:233-250): level i holds 2^(nv-i) G1 points.

    `levels` is a list of (n_i, 64) uint8 arrays; level 0 alone is enough for commit / batch_commit.
    The reference sketches exactly this registration as `gpu_g1` (pst13.rs:52-61)."""

    def __init__(self, ctx, levels):
        self.ctx = ctx
        self.level_srs = [ctx.srs_register(lv) for lv in levels]
        self.num_vars = int(np.log2(levels[0].shape[0]))

    @property
    def srs(self):
        return self.level_srs[0]

    def release(self):
        for h in self.level_srs:
            self.ctx.srs_release(h)
        self.level_srs = []


def _ptr_array(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def commit(setup, evals, stride=32, form=0, max_num_bits=0):
    """PST13::commit (pst13.rs:282-296)."""
    evals = np.ascontiguousarray(evals, dtype=np.uint8)
    n = evals.size // stride
    out = np.zeros(COMMITMENT_BYTES, dtype=np.uint8)
    _check(_lib().cozk_pst13_commit(setup.ctx.handle, setup.srs, evals.ctypes.data_as(ctypes.c_void_p), n, stride, form,
                                    max_num_bits, out.ctypes.data_as(ctypes.c_void_p)))
    return PST13Commitment.from_bytes(out)


def batch_commit(setup, polys, stride=32, form=0, max_num_bits=None):
    """PST13::batch_commit (pst13.rs:299-331): equal lengths asserted, "Key length error" when longer than the SRS."""
    polys = [np.ascontiguousarray(p, dtype=np.uint8) for p in polys]
    n = polys[0].size // stride
    assert all(p.size // stride == n for p in polys), "batch commit requires all batches to have the same length"
    k = len(polys)
    out = np.zeros((k, COMMITMENT_BYTES), dtype=np.uint8)
    bits = (ctypes.c_uint * k)(*max_num_bits) if max_num_bits is not None else None
    _check(_lib().cozk_pst13_batch_commit(setup.ctx.handle, setup.srs, _ptr_array(polys), k, n, stride, form, bits,
                                          out.ctypes.data_as(ctypes.c_void_p)))
    return [PST13Commitment.from_bytes(out[j]) for j in range(k)]


def batch_commit_rep3(setup, polys, is_shared, commit_to_public, form=0, max_num_bits=None):
    """PST13::batch_commit_rep3 (pst13.rs:165-229).  Shared polynomials are (n, 64) AoS {a, b} share arrays (share a
    is read in place); public ones are (n, 32).  Returns a list with None for MaybeShared::Public(None)."""
    polys = [np.ascontiguousarray(p, dtype=np.uint8) for p in polys]
    k = len(polys)
    n = polys[0].size // (64 if is_shared[0] else 32)
    for p, sh in zip(polys, is_shared):  # pst13.rs:307-309 asserts equal lengths; a short buffer would be read past its end
        if p.size != n * (64 if sh else 32):
            raise ValueError("polynomial holds %d bytes, %d coefficients need %d" % (p.size, n, n * (64 if sh else 32)))
    flags = (ctypes.c_uint8 * k)(*[1 if s else 0 for s in is_shared])
    present = (ctypes.c_uint8 * k)()
    out = np.zeros((k, COMMITMENT_BYTES), dtype=np.uint8)
    bits = (ctypes.c_uint * k)(*max_num_bits) if max_num_bits is not None else None
    _check(_lib().cozk_pst13_batch_commit_rep3(setup.ctx.handle, setup.srs, _ptr_array(polys), flags, k, n, form, bits,
                                               1 if commit_to_public else 0, out.ctypes.data_as(ctypes.c_void_p), present))
    return [PST13Commitment.from_bytes(out[j]) if present[j] else None for j in range(k)]


ELEM_BYTES = {0: 64, 1: 32, 2: 1, 3: 2, 4: 4, 5: 8, 6: 8}  # COZK_POLY_SHARED, PUBLIC, U8, U16, U32, U64, I64


def batch_commit_packed(setup, polys, kinds, commit_to_public=True):
    """PST13::batch_commit / batch_commit_rep3 over the reference's packed in-memory polynomial forms
    (MultilinearPolynomial::{LargeScalars, U8Scalars .. I64Scalars}, multilinear_polynomial.rs:226-268, and
    Rep3DensePolynomial share arrays): small-scalar polynomials cross PCIe at 1 - 8 bytes per coefficient and are
    widened on the device.  polys[j]: numpy array in its natural dtype (uint8 .. int64) or (n, 32) / (n, 64) uint8 images;
    kinds[j]: a COZK_POLY_* constant (co-zkvms_b200.rep3: SHARED, PUBLIC, U8, U16, U32, U64, I64)."""
    arrs = [np.ascontiguousarray(p) for p in polys]
    k = len(arrs)
    n = arrs[0].nbytes // ELEM_BYTES[kinds[0]]
    for a, kd in zip(arrs, kinds):  # the reference asserts equal lengths (pst13.rs:307-309); the C ABI has no length arguments
        if a.nbytes != n * ELEM_BYTES[kd]:
            raise ValueError("polynomial of kind %d holds %d bytes, %d coefficients need %d" % (kd, a.nbytes, n, n * ELEM_BYTES[kd]))
    ptrs = (ctypes.c_void_p * k)(*[a.ctypes.data for a in arrs])
    kd = (ctypes.c_int * k)(*kinds)
    present = (ctypes.c_uint8 * k)()
    out = np.zeros((k, COMMITMENT_BYTES), dtype=np.uint8)
    _check(_lib().cozk_pst13_batch_commit_packed(setup.ctx.handle, setup.srs, ptrs, kd, k, n, 1 if commit_to_public else 0,
                                                 out.ctypes.data_as(ctypes.c_void_p), present))
    return [PST13Commitment.from_bytes(out[j]) if present[j] else None for j in range(k)]


def open(setup, evals, point, stride=32):  # noqa: A001 - the reference's name
    """open() behind PST13::prove_rep3 (pst13.rs:428-474).  Returns (proofs (nv, 72), evaluation (32,))."""
    evals = np.ascontiguousarray(evals, dtype=np.uint8)
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    nv = point.shape[0]
    if evals.size < ((1 << nv) - 1) * stride + 32:  # pst13.rs:438 asserts the polynomial size
        raise ValueError("2^%d evaluations at stride %d need %d bytes, got %d" % (nv, stride, ((1 << nv) - 1) * stride + 32, evals.size))
    srs = (ctypes.c_uint64 * nv)(*setup.level_srs[:nv])
    proofs = np.zeros((nv, 72), dtype=np.uint8)
    ev = np.zeros(32, dtype=np.uint8)
    _check(_lib().cozk_pst13_open(setup.ctx.handle, srs, nv, evals.ctypes.data_as(ctypes.c_void_p), stride,
                                  point.ctypes.data_as(ctypes.c_void_p), 0, proofs.ctypes.data_as(ctypes.c_void_p),
                                  ev.ctypes.data_as(ctypes.c_void_p)))
    return proofs, ev


def combine_commitment_shares(commitments):
    """PST13::combine_commitment_shares (pst13.rs:72-108) for the all-Shared case: c0 + c1 + c2."""
    raw = np.stack([c.to_bytes() for c in commitments])
    out = np.zeros(COMMITMENT_BYTES, dtype=np.uint8)
    _check(_lib().cozk_pst13_combine_commitment_shares(raw.ctypes.data_as(ctypes.c_void_p), len(commitments),
                                                       out.ctypes.data_as(ctypes.c_void_p)))
    return PST13Commitment.from_bytes(out)


def coordinate_prove(party_proofs):
    """PST13::coordinate_prove (pst13.rs:110-122): element-wise sum of the parties' proof vectors."""
    arr = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.uint8).reshape(-1, 72) for p in party_proofs]))
    parties, length = arr.shape[0], arr.shape[1]
    out = np.zeros((length, 72), dtype=np.uint8)
    _check(_lib().cozk_pst13_coordinate_prove(arr.ctypes.data_as(ctypes.c_void_p), parties, length,
                                              out.ctypes.data_as(ctypes.c_void_p)))
    return out


def combine_comm(commitments):
    """combine_comm (snarks-core/src/poly/commitment.rs:56-63): sum of chunk commitments, nv += log2(#chunks)."""
    raw = np.stack([c.to_bytes() for c in commitments])
    out = np.zeros(COMMITMENT_BYTES, dtype=np.uint8)
    _check(_lib().cozk_combine_comm(raw.ctypes.data_as(ctypes.c_void_p), len(commitments), out.ctypes.data_as(ctypes.c_void_p)))
    return PST13Commitment.from_bytes(out)
