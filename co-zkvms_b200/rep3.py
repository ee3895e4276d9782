"""Python face of include/cozk_rep3.h: device-resident Rep3 polynomials and the steps either side of the MSM.

Names follow the reference so that the parity tests read like its code:
  Rep3DensePolynomial            co-jolt/src/poly/dense_mlpoly.rs:23-32 (from_wire = receive_request + deserialize)
  linear_combination             co-jolt/src/poly/multilinear_polynomial.rs:196-296
  evaluate_at_chi/batch_evaluate co-jolt/src/poly/dense_mlpoly.rs:160-194 (chis from the caller, or eq_evals on the device)
  eq_evals                       jolt-core EqPolynomial::evals order (msb_first) / ark-poly evaluate order (lsb first)
  batch_commit_rep3              co-jolt/src/poly/commitment/pst13.rs:165-229 over resident polynomials
  prove_rep3                     pst13.rs:125-137 (opening point reversed, share a opened)
Field elements are 32-byte little-endian Montgomery values unless stated otherwise.
"""
import ctypes

import numpy as np

SHARED, PUBLIC, U8, U16, U32, U64, I64 = range(7)
ERR_WIRE = -6
_SMALL_DTYPES = {U8: np.uint8, U16: np.dtype("<u2"), U32: np.dtype("<u4"), U64: np.dtype("<u8"), I64: np.dtype("<i8")}


def _lib():
    from . import lib
    return lib()


def _check(rc):
    from . import _check as chk
    chk(rc)


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Rep3DensePolynomial:
    """A polynomial that lives in HBM (cozk_poly).  kind SHARED: Rep3 share pairs; otherwise a public polynomial."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    # ---- constructors
    @classmethod
    def upload(cls, ctx, coeffs, kind=SHARED, device=0):
        """coeffs: SHARED (n, 64) uint8 AoS {a, b}; PUBLIC (n, 32) uint8; small kinds: integer array of n values."""
        if kind in _SMALL_DTYPES:
            arr = np.ascontiguousarray(coeffs, dtype=_SMALL_DTYPES[kind])
            n = arr.size
        else:
            arr = np.ascontiguousarray(coeffs, dtype=np.uint8)
            n = arr.size // (64 if kind == SHARED else 32)
        h = ctypes.c_uint64()
        _check(_lib().cozk_poly_upload(ctx.handle, device, _vp(arr), n, kind, ctypes.byref(h)))
        return cls(ctx, h.value)

    @classmethod
    def from_device(cls, ctx, dbuf, n, kind=SHARED, device=0):
        """dbuf: a DeviceBuffer holding the in-memory image (copied device to device)."""
        h = ctypes.c_uint64()
        _check(_lib().cozk_poly_from_device(ctx.handle, device, ctypes.c_void_p(dbuf.ptr), n, kind, ctypes.byref(h)))
        return cls(ctx, h.value)

    @classmethod
    def from_wire(cls, ctx, raw, tagged=False, device=0):
        """raw: bytes of one ark-serialize (uncompressed) Rep3DensePolynomial.  Returns (polynomial, bytes consumed)."""
        buf = raw if isinstance(raw, np.ndarray) else np.frombuffer(bytes(raw), dtype=np.uint8)
        h, used = ctypes.c_uint64(), ctypes.c_size_t()
        _check(_lib().cozk_poly_from_wire(ctx.handle, device, _vp(buf), buf.size, 1 if tagged else 0, ctypes.byref(h),
                                          ctypes.byref(used)))
        return cls(ctx, h.value), used.value

    # ---- accessors
    def info(self):
        n, kind, dev = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_int()
        _check(_lib().cozk_poly_info(self.ctx.handle, self.handle, ctypes.byref(n), ctypes.byref(kind), ctypes.byref(dev)))
        return n.value, kind.value, dev.value

    def __len__(self):
        return self.info()[0]

    @property
    def kind(self):
        return self.info()[1]

    def download(self):
        n, kind, _ = self.info()
        out = np.zeros((n, 64 if kind == SHARED else 32), dtype=np.uint8)
        _check(_lib().cozk_poly_download(self.ctx.handle, self.handle, _vp(out)))
        return out

    def release(self):
        if self.handle:
            _check(_lib().cozk_poly_release(self.ctx.handle, self.handle))
            self.handle = 0

    # ---- the reference's methods
    def evaluate_at_chi(self, chis):
        return batch_evaluate_at_chi([self], chis)[0]


def _handles(polys):
    return (ctypes.c_uint64 * len(polys))(*[p.handle for p in polys])


def linear_combination(polynomials, coefficients, party_id):
    """Rep3MultilinearPolynomial::linear_combination: sum_j coefficients[j] * polynomials[j], device to device."""
    ctx = polynomials[0].ctx
    coeffs = np.ascontiguousarray(coefficients, dtype=np.uint8).reshape(len(polynomials), 32)
    h = ctypes.c_uint64()
    _check(_lib().cozk_rep3_linear_combination(ctx.handle, _handles(polynomials), _vp(coeffs), len(polynomials), int(party_id),
                                               ctypes.byref(h)))
    return Rep3DensePolynomial(ctx, h.value)


def batch_evaluate_at_chi(polys, chis):
    """[poly.evaluate_at_chi(chis) for poly in polys] (batch_evaluate with the eq table supplied): (k, 32) uint8.
    chis: host array of n Montgomery values, or a device-resident public polynomial (eq_evals)."""
    ctx = polys[0].ctx
    out = np.zeros((len(polys), 32), dtype=np.uint8)
    if isinstance(chis, Rep3DensePolynomial):
        _check(_lib().cozk_rep3_evaluate_at_chi_poly(ctx.handle, _handles(polys), len(polys), chis.handle, _vp(out)))
        return out
    chis = np.ascontiguousarray(chis, dtype=np.uint8).reshape(-1, 32)
    _check(_lib().cozk_rep3_evaluate_at_chi(ctx.handle, _handles(polys), len(polys), _vp(chis), chis.shape[0], _vp(out)))
    return out


def eq_evals(ctx, point, msb_first=True, device=0):
    """The eq table of `point` ((nv, 32) Montgomery) as a device-resident public polynomial of 2^nv values.
    msb_first=True: jolt-core's EqPolynomial::evals(r) (point[0] pairs with the top index bit) - the chis of
    evaluate_at_chi; False: the order ark-poly's evaluate / open() use (point[i] pairs with index bit i)."""
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    h = ctypes.c_uint64()
    _check(_lib().cozk_eq_evals(ctx.handle, device, _vp(point), point.shape[0], 1 if msb_first else 0, ctypes.byref(h)))
    return Rep3DensePolynomial(ctx, h.value)


def batch_commit_rep3(setup, polys, commit_to_public):
    """PST13::batch_commit_rep3 over resident polynomials; None marks MaybeShared::Public(None)."""
    from .pst13 import COMMITMENT_BYTES, PST13Commitment
    k = len(polys)
    out = np.zeros((k, COMMITMENT_BYTES), dtype=np.uint8)
    present = (ctypes.c_uint8 * k)()
    _check(_lib().cozk_pst13_batch_commit_polys(setup.ctx.handle, setup.srs, _handles(polys), k, 1 if commit_to_public else 0,
                                                _vp(out), present))
    return [PST13Commitment.from_bytes(out[j]) if present[j] else None for j in range(k)]


def pair_sums(ctx, srs):
    h = ctypes.c_uint64()
    _check(_lib().cozk_srs_pair_sums(ctx.handle, srs, ctypes.byref(h)))
    return h.value


def create_open_key(setup, nv=None):
    """Adds `open_key` to a PST13Setup: pair sums of every level, small levels concatenated (setup-time work)."""
    nv = len(setup.level_srs) if nv is None else nv
    srs = (ctypes.c_uint64 * nv)(*setup.level_srs[:nv])
    h = ctypes.c_uint64()
    _check(_lib().cozk_pst13_open_key_create(setup.ctx.handle, srs, nv, ctypes.byref(h)))
    setup.open_key = h.value
    return setup


def release_open_key(setup):
    if getattr(setup, "open_key", 0):
        _check(_lib().cozk_pst13_open_key_release(setup.ctx.handle, setup.open_key))
    setup.open_key = 0


def prove_rep3(setup, poly, opening_point, keyed=True):
    """PST13::prove_rep3 (pst13.rs:125-137) without the network send: the opening point is reversed, share a is opened.
    Returns (proofs (nv, 72), evaluation (32,))."""
    point = np.ascontiguousarray(opening_point, dtype=np.uint8).reshape(-1, 32)[::-1].copy()
    return open_poly(setup, poly, point, keyed=keyed)


def open_poly(setup, poly, point, keyed=True):
    """open() (pst13.rs:428-474) on a resident polynomial; `point` in the order open() receives it.  keyed: use the
    setup's opening key (pair sums + batched small levels) when it has one, else the reference's schedule."""
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    nv = point.shape[0]
    srs = (ctypes.c_uint64 * nv)(*setup.level_srs[:nv])
    key = getattr(setup, "open_key", 0) if keyed else 0
    proofs = np.zeros((nv, 72), dtype=np.uint8)
    ev = np.zeros(32, dtype=np.uint8)
    _check(_lib().cozk_pst13_open_poly(setup.ctx.handle, srs, nv, key, poly.handle, _vp(point), _vp(proofs), _vp(ev)))
    return proofs, ev


def open_keyed(setup, evals, point, stride=32):
    """cozk_pst13_open_keyed: host evaluations, the setup's opening key."""
    evals = np.ascontiguousarray(evals, dtype=np.uint8)
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    nv = point.shape[0]
    proofs = np.zeros((nv, 72), dtype=np.uint8)
    ev = np.zeros(32, dtype=np.uint8)
    _check(_lib().cozk_pst13_open_keyed(setup.ctx.handle, setup.open_key, _vp(evals), stride, _vp(point), 0, _vp(proofs), _vp(ev)))
    return proofs, ev


def last_stats(ctx):
    s = (ctypes.c_double * 8)()
    _check(_lib().cozk_rep3_last_stats(ctx.handle, s))
    return {"h2d_ms": s[0], "ingest_ms": s[1], "lincomb_ms": s[2], "chi_ms": s[3], "lincomb_bytes": s[4],
            "partial_sum_ms": s[5], "peer_bytes": s[6]}
