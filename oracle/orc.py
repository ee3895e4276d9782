"""ctypes binding of the CPU oracle (oracle/bn254.c).  TEST INFRASTRUCTURE ONLY - see oracle/bn254.h.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product package.  All buffers are numpy uint8/uint64 arrays in the C ABI's wire formats.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")

MONT, CANON = 0, 1
DIST = {"uniform": 0, "const": 1, "wminus": 2, "dup": 3, "small16": 4, "zero_half": 5}


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("bn254.c", "bn254.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp, sz, u64, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
        L.orc_field_op.argtypes = [ci, ci, vp, vp, vp, sz]
        L.orc_g1_op.argtypes = [ci, vp, vp, vp, sz]
        L.orc_g1_is_valid.argtypes = [vp]
        L.orc_g1_is_valid.restype = ci
        L.orc_g1_mul.argtypes = [vp, vp, ci, vp]
        L.orc_gen_bases.argtypes = [u64, sz, sz, vp, ci]
        L.orc_gen_scalars.argtypes = [ci, u64, sz, sz, sz, ci, vp, sz]
        L.orc_msm_naive.argtypes = [vp, vp, sz, ci, sz, vp]
        L.orc_msm.argtypes = [vp, vp, sz, ci, sz, ci, vp]
        L.orc_msm_window.argtypes = [sz]
        L.orc_msm_window.restype = ci
        L.orc_msm_last_counts.argtypes = [vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def ncores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def field_op(which, op, a, b=None):
    """which: 'fq'|'fr'; op: add sub mul sqr neg inv to_mont from_mont; arrays (n,4) uint64."""
    ops = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "neg": 4, "inv": 5, "to_mont": 6, "from_mont": 7}
    a = np.ascontiguousarray(a, dtype=np.uint64)
    bb = np.ascontiguousarray(b, dtype=np.uint64) if b is not None else None
    out = np.empty_like(a)
    lib().orc_field_op(0 if which == "fq" else 1, ops[op], _p(a), _p(bb), _p(out), a.shape[0])
    return out


def g1_op(op, a, b=None):
    """op: add dbl neg on (n,72) uint8 arrays of wire points."""
    ops = {"add": 0, "dbl": 1, "neg": 2}
    a = np.ascontiguousarray(a, dtype=np.uint8)
    bb = np.ascontiguousarray(b, dtype=np.uint8) if b is not None else None
    out = np.empty_like(a)
    lib().orc_g1_op(ops[op], _p(a), _p(bb), _p(out), a.shape[0])
    return out


def g1_is_valid(pt72):
    pt72 = np.ascontiguousarray(pt72, dtype=np.uint8)
    return bool(lib().orc_g1_is_valid(_p(pt72)))


def g1_mul(pt72, scalar32, form=CANON):
    pt72 = np.ascontiguousarray(pt72, dtype=np.uint8)
    s = np.ascontiguousarray(scalar32, dtype=np.uint8)
    out = np.zeros(72, dtype=np.uint8)
    lib().orc_g1_mul(_p(pt72), _p(s), form, _p(out))
    return out


def gen_bases(seed, n, start=0, threads=None):
    out = np.empty((n, 64), dtype=np.uint8)
    lib().orc_gen_bases(seed, start, n, _p(out), threads or ncores())
    return out


def gen_scalars(dist, seed, n, form=MONT, stride=32, start=0, total_n=None):
    out = np.zeros((n, stride), dtype=np.uint8)
    lib().orc_gen_scalars(DIST[dist], seed, start, n, total_n if total_n is not None else n, form, _p(out), stride)
    return out


def msm(bases, scalars, form=MONT, threads=None, n=None):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    n = min(bases.shape[0], scalars.shape[0]) if n is None else n
    out = np.zeros(72, dtype=np.uint8)
    lib().orc_msm(_p(bases), _p(scalars), scalars.shape[1], form, n, threads or ncores(), _p(out))
    return out


def msm_naive(bases, scalars, form=MONT, n=None):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    n = min(bases.shape[0], scalars.shape[0]) if n is None else n
    out = np.zeros(72, dtype=np.uint8)
    lib().orc_msm_naive(_p(bases), _p(scalars), scalars.shape[1], form, n, _p(out))
    return out


def msm_window(n):
    return lib().orc_msm_window(n)


def last_counts():
    a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    lib().orc_msm_last_counts(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


# ---- conversions between wire bytes and Python ints (for comparison with oracle/pyref.py)

def int_to_le32(x):
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint8)


def le32_to_int(b):
    return int.from_bytes(bytes(bytearray(b)), "little")


def point_to_wire(pt):
    """pyref affine point (canonical ints) or None -> 72-byte wire point (Montgomery)."""
    from . import pyref
    out = np.zeros(72, dtype=np.uint8)
    if pt is None:
        out[64] = 1
        return out
    out[0:32] = int_to_le32(pyref.to_mont(pt[0], pyref.P))
    out[32:64] = int_to_le32(pyref.to_mont(pt[1], pyref.P))
    return out


def wire_to_point(w):
    from . import pyref
    w = np.asarray(w, dtype=np.uint8)
    if len(w) >= 72 and w[64]:
        return None
    return (pyref.from_mont(le32_to_int(w[0:32]), pyref.P), pyref.from_mont(le32_to_int(w[32:64]), pyref.P))
