"""Host emulation of the kernel thread bodies (test harness only; see emul.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "co-zkvms_b200", "csrc")
LIB = os.path.join(HERE, "_build", "libemul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [os.path.join(HERE, "emul.cpp")] + [os.path.join(CSRC, f) for f in
                                                   ("field.cuh", "curve.cuh", "msm_kernels.cuh", "affine_kernels.cuh", "msm_plan.hpp", "rep3_kernels.cuh")]
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB,
                                   os.path.join(HERE, "emul.cpp")])
        L = ctypes.CDLL(LIB)
        vp, sz, u32, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int
        L.emul_msm.argtypes = [vp, sz, vp, sz, sz, ci, u32, u32, u32, vp, vp, vp, u32, sz, sz, u32]
        L.emul_choose_window.argtypes = [sz, u32, u32, sz, ctypes.c_double, vp]
        L.emul_choose_window.restype = None
        L.emul_set_dominant.argtypes = [ci]
        L.emul_set_dominant.restype = None
        L.emul_build_table.argtypes = [vp, sz, u32, u32, ci, sz, vp]
        L.emul_build_table.restype = None
        L.emul_set_reduce_2d.argtypes = [ci]
        L.emul_set_reduce_2d.restype = None
        L.emul_set_affine_rounds.argtypes = [ci]
        L.emul_set_affine_rounds.restype = None
        L.emul_affine_stats.argtypes = [vp]
        L.emul_affine_stats.restype = None
        L.emul_affine_round.argtypes = [vp, vp, sz, vp, vp, vp, vp, vp, vp, vp, u32]
        L.emul_affine_round.restype = None
        L.emul_set_acc_chunk.argtypes = [sz, ci]
        L.emul_set_acc_chunk.restype = None
        L.emul_field_op.argtypes = [ci, vp, vp, vp, sz]
        L.emul_g1_op.argtypes = [ci, vp, vp, vp, sz]
        L.emul_ingest.argtypes = [vp, sz, vp]
        L.emul_widen.argtypes = [vp, u32, u32, sz, vp]
        L.emul_lincomb.argtypes = [vp, vp, vp, u32, vp, u32, u32, sz, vp]
        L.emul_chi.argtypes = [vp, vp, vp, u32, vp, sz, u32, vp]
        L.emul_sum_partials.argtypes = [vp, vp, vp, u32, u32, u32, sz, vp]
        L.emul_pair_sum.argtypes = [vp, vp, sz, vp, vp]
        L.emul_eq.argtypes = [vp, u32, u32, ci, vp]
        L.emul_wide_dot.argtypes = [vp, vp, sz, u32, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def choose_window(n, g, bits, max_buckets, poison=0.0):
    """(window size or 0 when none fits, largest group size <= g that fits) as the engine's plan chooses them."""
    out = np.zeros(2, np.uint32)
    lib().emul_choose_window(n, g, bits, max_buckets, poison, _p(out))
    return int(out[0]), int(out[1])


def set_dominant(on):
    """1: calls that cover the whole SRS go through the engine's dominant-digit path (analysis pass, compacted segments,
    row totals); 0: the plain pair layout."""
    lib().emul_set_dominant(1 if on else 0)


def build_table(bases, c, rows, mode=0, slab=None):
    """The SRS table (rows x n affine points, row w = 2^(c w) P): mode 0 = the row-by-row contract body, 1 = the engine's two
    steps (chain in XYZZ, one inversion per point) in slabs of `slab` points."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    n = bases.shape[0]
    out = np.zeros((rows * n, 64), np.uint8)
    lib().emul_build_table(_p(bases), n, c, rows, mode, slab or n, _p(out))
    return out


def set_reduce_2d(on):
    """1: the bucket reduce runs in its row / column form (two tree sums, MsmPlan::reduce_2d); 0: group sums + masked sums."""
    lib().emul_set_reduce_2d(1 if on else 0)


def set_affine_rounds(rounds):
    """> 0: the batched-affine pre-reduction runs in front of the accumulate levels, with the engine's rule for how many rounds
    pay (msm.cu, affine_rounds_for); < 0: exactly -rounds rounds whatever the run lengths; 0: off."""
    lib().emul_set_affine_rounds(rounds)


def affine_stats():
    """(entries of the reduced list, overflow entries of all rounds) of the last chunk of the last msm() call."""
    out = np.zeros(2, np.uint32)
    lib().emul_affine_stats(_p(out))
    return int(out[0]), int(out[1])


def affine_round(keys, vals, pts, cap=None):
    """One halving round of the contract body: ((keys, vals, points), (overflow keys, overflow points))."""
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    vals = np.ascontiguousarray(vals, dtype=np.uint32)
    pts = np.ascontiguousarray(pts, dtype=np.uint8)
    m = keys.size
    mo = (m + 1) // 2
    cap = mo if cap is None else cap
    ko, vo, po = np.zeros(mo, np.uint32), np.zeros(mo, np.uint32), np.zeros((mo, 64), np.uint8)
    cnt, ok, op = np.zeros(1, np.uint32), np.zeros(max(cap, 1), np.uint32), np.zeros((max(cap, 1), 64), np.uint8)
    lib().emul_affine_round(_p(keys), _p(vals), m, _p(pts), _p(ko), _p(vo), _p(po), _p(cnt), _p(ok), _p(op), cap)
    c = int(cnt[0])
    return (ko, vo, po), (ok[:c], op[:c])


def set_acc_chunk(resident=0, force_l=0):
    """Level-1 chunk length of the accumulate stage: `resident` threads per wave (the plan then picks the length that
    fills the last wave, as the engine does with SMs x 512), or a fixed `force_l`; (0, 0) = the default of 32."""
    lib().emul_set_acc_chunk(resident, force_l)


def msm(bases, scalars, form=0, g=1, bits=0, c=0, stride=32, infinity=None, table_c=0, n=None, base_offset=0,
        stream_chunks=1):
    """scalars: (g*n, stride) uint8 laid out vector after vector.  table_c != 0: `bases` is the whole registered SRS,
    a table of 2^(table_c*w) * P rows is built from it and the call covers [base_offset, base_offset + n)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    srs_n = bases.shape[0]
    n = srs_n if n is None else n
    out = np.zeros((g, 72), np.uint8)
    st = np.zeros(6, np.uint32)  # window, windows, accumulate levels, sum chunks, pairs, dominant-digit path taken
    inf = np.ascontiguousarray(infinity, dtype=np.uint8) if infinity is not None else None
    rc = lib().emul_msm(_p(bases), n, _p(scalars), n * stride, stride, form, g, bits, c, _p(inf), _p(out), _p(st),
                        table_c, srs_n, base_offset, stream_chunks)
    assert rc == 0, rc
    return out, st


def field_op(op, a, b=None):
    ops = {"mul": 0, "add": 1, "sub": 2, "sqr": 3, "inv": 4, "fr_from_mont": 5,
           "f29_mul": 10, "f29_add": 11, "f29_sub": 12, "f29_sqr": 13, "f29_inv": 14}
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
    bb = np.ascontiguousarray(b, dtype=np.uint8) if b is not None else None
    out = np.zeros_like(a)
    lib().emul_field_op(ops[op], _p(a), _p(bb), _p(out), a.shape[0])
    return out


def g1_op(op, a, b=None):
    ops = {"add": 0, "madd": 1, "dbl": 2, "f29_madd3": 3}
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 72)
    bb = np.ascontiguousarray(b, dtype=np.uint8) if b is not None else None
    out = np.zeros_like(a)
    lib().emul_g1_op(ops[op], _p(a), _p(bb), _p(out), a.shape[0])
    return out


# ---- rep3_kernels.cuh bodies (rows N1, N2, N4)
KIND = {"shared": 0, "mont": 1, "canon": 2}


def ingest(data):
    """data: (n, 32) canonical little-endian -> Montgomery, in place.  Returns the `bad` flag."""
    bad = np.zeros(1, np.uint32)
    lib().emul_ingest(_p(data), data.shape[0], _p(bad))
    return int(bad[0])


def widen(raw, elem_bytes, signed=False):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    n = raw.size // elem_bytes
    out = np.zeros((n, 32), np.uint8)
    lib().emul_widen(_p(raw), elem_bytes, 1 if signed else 0, n, _p(out))
    return out


def _descs(polys):
    arrs = [np.ascontiguousarray(a, dtype=np.uint8) for _, a in polys]
    ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    lens = np.array([a.shape[0] for a in arrs], np.uint64)
    kinds = np.array([KIND[k] for k, _ in polys], np.uint32)
    return arrs, ptrs, lens, kinds


def lincomb(polys, coeffs, party):
    """polys: list of (kind, array): "shared" (n, 64) Montgomery AoS, "mont" (n, 32), "canon" (n, 32)."""
    arrs, ptrs, lens, kinds = _descs(polys)
    shared_out = 1 if any(k == "shared" for k, _ in polys) else 0
    n = int(lens.max())
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint8)
    out = np.zeros((n, 64 if shared_out else 32), np.uint8)
    lib().emul_lincomb(ptrs, _p(lens), _p(kinds), len(arrs), _p(coeffs), party, shared_out, n, _p(out))
    return out


def lincomb_by_device(polys, coeffs, party, device_of):
    """The multi-device linear combination as the engine runs it: one partial combination per device (a partial is
    shared when its device holds a shared polynomial, public otherwise), then the summing body.  device_of[j]: device of
    polynomial j."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint8).reshape(len(polys), 32)
    any_shared = any(k == "shared" for k, _ in polys)
    parts = []
    for dev in sorted(set(device_of)):
        idx = [j for j, d in enumerate(device_of) if d == dev]
        sub = [polys[j] for j in idx]
        out = lincomb(sub, coeffs[idx], party)
        parts.append(("shared" if any(k == "shared" for k, _ in sub) else "mont", out))
    arrs, ptrs, lens, kinds = _descs(parts)
    n = int(lens.max())
    out = np.zeros((n, 64 if any_shared else 32), np.uint8)
    lib().emul_sum_partials(ptrs, _p(lens), _p(kinds), len(arrs), party, 1 if any_shared else 0, n, _p(out))
    return out


def chi(polys, chis, T=32):
    arrs, ptrs, lens, kinds = _descs(polys)
    chis = np.ascontiguousarray(chis, dtype=np.uint8)
    out = np.zeros((len(arrs), 32), np.uint8)
    lib().emul_chi(ptrs, _p(lens), _p(kinds), len(arrs), _p(chis), chis.shape[0], T, _p(out))
    return out


def pair_sum(bases, infinity=None):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    half = bases.shape[0] // 2
    inf = np.ascontiguousarray(infinity, dtype=np.uint8) if infinity is not None else None
    out, oi = np.zeros((half, 64), np.uint8), np.zeros(half, np.uint8)
    lib().emul_pair_sum(_p(bases), _p(inf), half, _p(out), _p(oi))
    return out, oi


def wide_dot(a, b, repeat=1):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros(32, np.uint8)
    lib().emul_wide_dot(_p(a), _p(b), a.shape[0], repeat, _p(out))
    return out


def eq(point, msb_first=False, lo_bits=None):
    """point: (nv, 32) Montgomery -> (2^nv, 32) eq table through the two thread bodies of cozk_eq_evals."""
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    nv = point.shape[0]
    lo_bits = nv // 2 if lo_bits is None else lo_bits
    out = np.zeros((1 << nv, 32), np.uint8)
    lib().emul_eq(_p(point), nv, lo_bits, 1 if msb_first else 0, _p(out))
    return out
