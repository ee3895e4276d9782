"""co-spartan's worker-side commitment functions over the C ABI (host-side mirror; names follow the reference).

  poly_commit_worker                  co-noir-spartan/co-spartan/src/worker.rs:577-590   MultilinearPC::commit per polynomial
  distributed_open                    worker.rs:774-809                                  nv folds + nv msm_bigint calls
  aggregate_poly                      co-noir-spartan/co-spartan/src/utils.rs:85-107     sum_j eta^j * polys[j]
  distributed_batch_open_poly_worker  worker.rs:745-772                                  aggregate + open + evaluate, no network send
  combine_comm                        snarks-core/src/poly/commitment.rs:56-63           sum of the workers' chunk commitments

Polynomials are the party's `share_0.evaluations` (mpc-core/src/protocols/rep3/poly.rs:10-14): dense vectors of Fr, 32-byte
little-endian Montgomery values, kept on the device as public polynomial handles.  `ck` is a pst13.PST13Setup (the
CommitterKey's powers_of_g levels registered as SRS handles); give it an opening key (rep3.create_open_key) to get the
pair-sum schedule.
"""
import ctypes

import numpy as np

from . import rep3


def _lib():
    from . import lib
    return lib()


def _check(rc):
    from . import _check as chk
    chk(rc)


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_MONT_R = (1 << 256) % _R_ORDER


def _fr_powers(eta, count):
    """[eta^0 .. eta^(count-1)] as (count, 32) Montgomery values (host integers: the `x *= eta` of aggregate_poly)."""
    e = int.from_bytes(bytes(np.ascontiguousarray(eta, dtype=np.uint8).reshape(32)), "little") * pow(_MONT_R, -1, _R_ORDER) % _R_ORDER
    out = np.zeros((count, 32), dtype=np.uint8)
    x = 1
    for j in range(count):
        out[j] = np.frombuffer((x * _MONT_R % _R_ORDER).to_bytes(32, "little"), dtype=np.uint8)
        x = x * e % _R_ORDER
    return out


def upload_evaluations(ctx, evaluations, device=0):
    """DenseMultilinearExtension.evaluations ((n, 32) Montgomery) -> device-resident polynomial."""
    return rep3.Rep3DensePolynomial.upload(ctx, evaluations, kind=rep3.PUBLIC, device=device)


def poly_commit_worker(ck, polys):
    """[MultilinearPC::commit(ck, p) for p in polys] (worker.rs:577-590): one batched MSM over the resident polynomials.
    Returns PST13Commitment-like objects (nv, g_product)."""
    return rep3.batch_commit_rep3(ck, list(polys), commit_to_public=True)


def aggregate_poly(eta, polys):
    """sum_j eta^j * polys[j] as a new device-resident polynomial (utils.rs:85-107).  eta: 32-byte Montgomery value."""
    return rep3.linear_combination(list(polys), _fr_powers(eta, len(polys)), 0)


def distributed_open(ck, polynomial, point, keyed=True):
    """(Proof.proofs (nv, 72), r[0][0] (32,)) (worker.rs:774-809) for a device-resident polynomial."""
    return rep3.open_poly(ck, polynomial, point, keyed=keyed)


def distributed_batch_open_poly_worker(polys, ck, point, eta, num_comms, keyed=True):
    """PartialProof{proofs, val, evals} (worker.rs:745-772) as a dict of arrays; one C call, everything on device 0."""
    polys = list(polys)
    ctx = ck.ctx
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    eta = np.ascontiguousarray(eta, dtype=np.uint8).reshape(32)
    nv = point.shape[0]
    srs = (ctypes.c_uint64 * nv)(*ck.level_srs[:nv])
    key = getattr(ck, "open_key", 0) if keyed else 0
    handles = (ctypes.c_uint64 * len(polys))(*[p.handle for p in polys])
    proofs = np.zeros((nv, 72), dtype=np.uint8)
    val = np.zeros(32, dtype=np.uint8)
    evals = np.zeros((len(polys), 32), dtype=np.uint8)
    _check(_lib().cozk_spartan_batch_open_worker(ctx.handle, key, srs, nv, handles, len(polys), num_comms, _vp(point), _vp(eta),
                                                 _vp(proofs), _vp(val), _vp(evals)))
    return {"proofs": proofs, "val": val, "evals": evals}


def combine_comm(commitments72):
    """combine_comm (snarks-core/src/poly/commitment.rs:56-63): the sum of the workers' chunk commitments."""
    from . import g1_sum
    return g1_sum(commitments72)
