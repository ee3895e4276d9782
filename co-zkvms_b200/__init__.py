"""co-zkvms_b200 - B200-native engine for the party-local BN254 G1 MSM of ChainSafe/co-zkvms.

This package is a thin ctypes binding of libcozk_msm.so (C ABI: include/cozk_msm.h; CUDA sources: csrc/).
The directory name carries a hyphen, so import it with importlib:

    cozk = importlib.import_module("co-zkvms_b200")

There is no CPU path: loading fails loudly when the library has not been built, and `Context()` fails when
no B200 is visible.  The oracle under oracle/ is never imported from here.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# COZK_LIB: load another build of the same library (kernel experiments measured side by side with tools/sweep.py)
LIB_PATH = os.environ.get("COZK_LIB") or os.path.join(HERE, "libcozk_msm.so")
TEST_LIB_PATH = os.path.join(HERE, "libcozk_test.so")  # include/cozk_test.h: generators, test kernels, microbenchmarks
SOURCES = ["msm.cu", "sort.cu", "affine.cu", "depth_kernels.cu", "aux.cu", "pst13.cu", "fixed_base.cu", "rep3poly.cu"]
TEST_SOURCES = ["testlib.cu"]
HEADERS = ["field.cuh", "field_ptx.inc", "curve.cuh", "msm_kernels.cuh", "sort_kernels.cuh", "affine_kernels.cuh", "rep3_kernels.cuh", "bulk_copy.cuh", "msm_plan.hpp", "engine.hpp",
           "depth_kernels.hpp"]
PUBLIC_HEADERS = ["cozk_msm.h", "cozk_rep3.h", "cozk_pst13.h", "cozk_test.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]

MONT, CANON = 0, 1
OK, ERR_INVALID_ARG, ERR_KEY_LENGTH, ERR_CUDA, ERR_NO_DEVICE, ERR_BAD_HANDLE, ERR_WIRE = 0, -1, -2, -3, -4, -5, -6
DIST = {"uniform": 0, "const": 1, "wminus": 2, "dup": 3, "small16": 4, "zero_half": 5}

# every symbol include/cozk_msm.h, cozk_pst13.h and cozk_test.h declare (checked by tests/test_abi.py without a GPU)
ABI_SYMBOLS = [
    "cozk_init", "cozk_destroy", "cozk_device_count", "cozk_srs_register", "cozk_srs_register_sliced", "cozk_srs_release", "cozk_last_stats_device", "cozk_srs_len",
    "cozk_msm_batch", "cozk_msm_batch_device", "cozk_msm_ragged_device", "cozk_g1_sum", "cozk_set_option", "cozk_last_stats", "cozk_last_error",
    "cozk_dev_alloc", "cozk_dev_free", "cozk_dev_upload", "cozk_dev_download", "cozk_host_alloc_pinned",
    "cozk_host_free_pinned", "cozk_dev_flush_l2", "cozk_srs_register_device",
    "cozk_pst13_commit", "cozk_pst13_batch_commit", "cozk_pst13_batch_commit_rep3", "cozk_pst13_batch_commit_packed", "cozk_pst13_open",
    "cozk_pst13_combine_commitment_shares", "cozk_pst13_coordinate_prove", "cozk_combine_comm",
    "cozk_fixed_base_batch_mul",
    # include/cozk_rep3.h
    "cozk_poly_upload", "cozk_poly_from_device", "cozk_poly_from_wire", "cozk_poly_release", "cozk_poly_info", "cozk_poly_download",
    "cozk_pst13_batch_commit_polys", "cozk_rep3_linear_combination", "cozk_rep3_evaluate_at_chi", "cozk_srs_pair_sums",
    "cozk_pst13_open_key_create", "cozk_pst13_open_key_release", "cozk_pst13_open_poly", "cozk_pst13_open_keyed",
    "cozk_rep3_last_stats", "cozk_rep3_evaluate_at_chi_poly", "cozk_eq_evals", "cozk_spartan_batch_open_worker",
]


# include/cozk_test.h, exported by libcozk_test.so
TEST_ABI_SYMBOLS = ["cozk_testgen_bases", "cozk_testgen_scalars", "cozk_test_field_op", "cozk_test_g1_op", "cozk_microbench",
                    "cozk_test_sort", "cozk_test_affine_rounds"]


class CozkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cozk error %d: %s" % (code, msg))
        self.code = code


def build(force=False, verbose=False):
    """Compile libcozk_msm.so for sm_100a with nvcc (cross-compiles without a GPU).  One object per translation unit
    under csrc/_obj/, stale ones rebuilt side by side, then one link."""
    from concurrent.futures import ThreadPoolExecutor
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(HERE, "..", "include", h) for h in PUBLIC_HEADERS]
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    jobs, objs, test_objs = [], [], []
    for s, into in [(x, objs) for x in SOURCES] + [(x, test_objs) for x in TEST_SOURCES]:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s + ".o")
        into.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_time, os.path.getmtime(src)):
            jobs.append([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(lambda c: subprocess.check_call(c, cwd=CSRC), jobs))
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    if not os.path.exists(LIB_PATH) or any(os.path.getmtime(LIB_PATH) < os.path.getmtime(o) for o in objs):
        subprocess.check_call([nvcc, "-shared"] + arch + ["-o", LIB_PATH] + objs, cwd=CSRC)
    # the test library calls into the product library (engine internals); the product library knows nothing of it
    if (not os.path.exists(TEST_LIB_PATH) or os.path.getmtime(TEST_LIB_PATH) < os.path.getmtime(LIB_PATH)
            or any(os.path.getmtime(TEST_LIB_PATH) < os.path.getmtime(o) for o in test_objs)):
        subprocess.check_call([nvcc, "-shared"] + arch + ["-o", TEST_LIB_PATH] + test_objs +
                              ["-L" + HERE, "-lcozk_msm", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"], cwd=CSRC)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libcozk_msm.so is not built (run __graft_entry__.build()); there is no fallback path")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, u64, ci, cu, cd = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint,
                               ctypes.POINTER(ctypes.c_double))
    pp = ctypes.POINTER(vp)
    L.cozk_init.argtypes = [pp, ctypes.POINTER(ci), ci]
    L.cozk_destroy.argtypes = [vp]
    L.cozk_destroy.restype = None
    L.cozk_device_count.argtypes = [vp]
    L.cozk_srs_register.argtypes = [vp, vp, sz, sz, vp, ctypes.POINTER(u64)]
    L.cozk_srs_register_sliced.argtypes = [vp, vp, sz, sz, vp, ctypes.POINTER(u64)]
    L.cozk_last_stats_device.argtypes = [vp, ci, cd]
    L.cozk_srs_register_device.argtypes = [vp, ci, vp, sz, ctypes.POINTER(u64)]
    L.cozk_srs_release.argtypes = [vp, u64]
    L.cozk_srs_len.argtypes = [vp, u64, ctypes.POINTER(sz)]
    L.cozk_msm_batch.argtypes = [vp, u64, sz, sz, pp, sz, sz, ci, cu, vp]
    L.cozk_msm_batch_device.argtypes = [vp, ci, u64, sz, sz, pp, sz, sz, ci, cu, vp]
    L.cozk_msm_ragged_device.argtypes = [vp, ci, u64, ctypes.POINTER(sz), ctypes.POINTER(sz), pp, sz, sz, ci, vp]
    L.cozk_g1_sum.argtypes = [vp, sz, vp]
    L.cozk_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_long]
    L.cozk_last_stats.argtypes = [vp, cd]
    L.cozk_last_error.restype = ctypes.c_char_p
    L.cozk_dev_alloc.argtypes = [vp, ci, sz, pp]
    L.cozk_dev_free.argtypes = [vp, ci, vp]
    L.cozk_dev_upload.argtypes = [vp, ci, vp, vp, sz]
    L.cozk_dev_download.argtypes = [vp, ci, vp, vp, sz]
    L.cozk_host_alloc_pinned.argtypes = [sz, pp]
    L.cozk_host_free_pinned.argtypes = [vp]
    L.cozk_dev_flush_l2.argtypes = [vp, ci]
    L.cozk_pst13_commit.argtypes = [vp, u64, vp, sz, sz, ci, cu, vp]
    L.cozk_pst13_batch_commit.argtypes = [vp, u64, pp, sz, sz, sz, ci, ctypes.POINTER(cu), vp]
    L.cozk_pst13_batch_commit_rep3.argtypes = [vp, u64, pp, ctypes.POINTER(ctypes.c_uint8), sz, sz, ci, ctypes.POINTER(cu), ci, vp,
                                               ctypes.POINTER(ctypes.c_uint8)]
    L.cozk_pst13_batch_commit_packed.argtypes = [vp, u64, pp, ctypes.POINTER(ci), sz, sz, ci, vp, ctypes.POINTER(ctypes.c_uint8)]
    L.cozk_pst13_open.argtypes = [vp, ctypes.POINTER(u64), sz, vp, sz, vp, ci, vp, vp]
    L.cozk_pst13_combine_commitment_shares.argtypes = [vp, sz, vp]
    L.cozk_pst13_coordinate_prove.argtypes = [vp, sz, sz, vp]
    L.cozk_combine_comm.argtypes = [vp, sz, vp]
    L.cozk_fixed_base_batch_mul.argtypes = [vp, vp, vp, sz, sz, ci, vp, ctypes.POINTER(u64)]
    pu64 = ctypes.POINTER(u64)
    L.cozk_poly_upload.argtypes = [vp, ci, vp, sz, ci, pu64]
    L.cozk_poly_from_device.argtypes = [vp, ci, vp, sz, ci, pu64]
    L.cozk_poly_from_wire.argtypes = [vp, ci, vp, sz, ci, pu64, ctypes.POINTER(sz)]
    L.cozk_poly_release.argtypes = [vp, u64]
    L.cozk_poly_info.argtypes = [vp, u64, ctypes.POINTER(sz), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.cozk_poly_download.argtypes = [vp, u64, vp]
    L.cozk_pst13_batch_commit_polys.argtypes = [vp, u64, pu64, sz, ci, vp, ctypes.POINTER(ctypes.c_uint8)]
    L.cozk_rep3_linear_combination.argtypes = [vp, pu64, vp, sz, ci, pu64]
    L.cozk_rep3_evaluate_at_chi.argtypes = [vp, pu64, sz, vp, sz, vp]
    L.cozk_srs_pair_sums.argtypes = [vp, u64, pu64]
    L.cozk_pst13_open_key_create.argtypes = [vp, pu64, sz, pu64]
    L.cozk_pst13_open_key_release.argtypes = [vp, u64]
    L.cozk_pst13_open_poly.argtypes = [vp, pu64, sz, u64, u64, vp, vp, vp]
    L.cozk_pst13_open_keyed.argtypes = [vp, u64, vp, sz, vp, ci, vp, vp]
    L.cozk_rep3_last_stats.argtypes = [vp, cd]
    L.cozk_rep3_evaluate_at_chi_poly.argtypes = [vp, pu64, sz, u64, vp]
    L.cozk_eq_evals.argtypes = [vp, ci, vp, sz, ci, pu64]
    L.cozk_spartan_batch_open_worker.argtypes = [vp, u64, pu64, sz, pu64, sz, sz, vp, vp, vp, vp, vp]
    _lib = L
    return L


_testlib = None


def testlib():
    """libcozk_test.so (include/cozk_test.h): synthetic inputs, test kernels, microbenchmarks - used by tests/, bench.py and
    tools/ only.  Loading it pulls in the product library first (it links against it)."""
    global _testlib
    if _testlib is not None:
        return _testlib
    lib()
    if not os.path.exists(TEST_LIB_PATH):
        raise ImportError("libcozk_test.so is not built (run __graft_entry__.build())")
    L = ctypes.CDLL(TEST_LIB_PATH)
    vp, sz, u64, ci, cu, cd = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint,
                               ctypes.POINTER(ctypes.c_double))
    L.cozk_testgen_bases.argtypes = [vp, ci, u64, sz, sz, vp]
    L.cozk_testgen_scalars.argtypes = [vp, ci, ci, u64, sz, sz, sz, ci, vp, sz]
    L.cozk_test_field_op.argtypes = [vp, ci, ci, vp, vp, vp, sz]
    L.cozk_test_g1_op.argtypes = [vp, ci, ci, vp, vp, vp, sz]
    L.cozk_test_sort.argtypes = [vp, ci, vp, vp, sz, cu, vp, sz, cu, sz, ci, cu, cu, sz, sz, ci, vp, vp]
    L.cozk_microbench.argtypes = [vp, ci, ci, ci, ci, ci, cd, cd]
    L.cozk_test_affine_rounds.argtypes = [vp, ci, vp, vp, sz, vp, sz, ci, ci, vp, vp, vp, ctypes.POINTER(sz), vp, vp, sz,
                                          ctypes.POINTER(cu), cd]
    _testlib = L
    return L


def _check(rc):
    if rc != 0:
        raise CozkError(rc, lib().cozk_last_error().decode("utf-8", "replace"))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(a)


class DeviceBuffer:
    """A raw device allocation owned by a Context."""

    def __init__(self, ctx, nbytes, device=0):
        self.ctx, self.nbytes, self.device = ctx, nbytes, device
        p = ctypes.c_void_p()
        _check(lib().cozk_dev_alloc(ctx.handle, device, nbytes, ctypes.byref(p)))
        self.ptr = p.value

    def upload(self, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        _check(lib().cozk_dev_upload(self.ctx.handle, self.device, ctypes.c_void_p(self.ptr + offset), _ptr(arr), arr.nbytes))
        return self

    def download(self, nbytes=None, offset=0):
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(nbytes, dtype=np.uint8)
        _check(lib().cozk_dev_download(self.ctx.handle, self.device, _ptr(out), ctypes.c_void_p(self.ptr + offset), nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().cozk_dev_free(self.ctx.handle, self.device, ctypes.c_void_p(self.ptr))
            self.ptr = None


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array."""

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        _check(lib().cozk_host_alloc_pinned(nbytes, ctypes.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            lib().cozk_host_free_pinned(ctypes.c_void_p(self.ptr))
            self.ptr = None


class Context:
    """Owns the devices of one process (cozk_ctx).  devices: list of CUDA device ids, default [0]."""

    def __init__(self, devices=None):
        L = lib()
        h = ctypes.c_void_p()
        if devices is None:
            rc = L.cozk_init(ctypes.byref(h), None, 1)
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = L.cozk_init(ctypes.byref(h), arr, len(devices))
        _check(rc)
        self.handle = h

    def close(self):
        if self.handle:
            lib().cozk_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def device_count(self):
        return lib().cozk_device_count(self.handle)

    # ---- SRS
    def srs_register(self, bases, infinity=None, stride=None, sliced=False):
        """bases: uint8 array (n, stride>=64), each row x||y Fq Montgomery (arkworks in-memory limbs).  sliced=True: every
        device of the context keeps only its point range (cozk_srs_register_sliced)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        n = bases.shape[0]
        stride = bases.shape[1] if stride is None else stride
        inf = np.ascontiguousarray(infinity, dtype=np.uint8) if infinity is not None else None
        if inf is not None and inf.size < n:
            raise ValueError("infinity flags: %d for %d bases" % (inf.size, n))
        h = ctypes.c_uint64()
        fn = lib().cozk_srs_register_sliced if sliced else lib().cozk_srs_register
        _check(fn(self.handle, _ptr(bases), n, stride, _ptr(inf), ctypes.byref(h)))
        return h.value

    def srs_register_device(self, dbuf, n, device=0):
        h = ctypes.c_uint64()
        _check(lib().cozk_srs_register_device(self.handle, device, ctypes.c_void_p(dbuf.ptr), n, ctypes.byref(h)))
        return h.value

    def srs_release(self, srs):
        _check(lib().cozk_srs_release(self.handle, srs))

    def srs_len(self, srs):
        n = ctypes.c_size_t()
        _check(lib().cozk_srs_len(self.handle, srs, ctypes.byref(n)))
        return n.value

    # ---- MSM
    def msm_batch(self, srs, scalars, n=None, base_offset=0, stride=32, form=MONT, max_num_bits=0):
        """scalars: list of uint8 arrays (each at least (n-1)*stride+32 bytes) or a 2-D/3-D array (k, n, stride).
        Returns (k, 72) uint8 wire points."""
        if isinstance(scalars, np.ndarray) and scalars.ndim == 3:
            vecs = [scalars[j] for j in range(scalars.shape[0])]
        elif isinstance(scalars, np.ndarray) and scalars.ndim == 2:
            vecs = [scalars]
        else:
            vecs = list(scalars)
        vecs = [np.ascontiguousarray(v, dtype=np.uint8) for v in vecs]
        k = len(vecs)
        if n is None:
            n = vecs[0].size // stride if k else 0
        for j, v in enumerate(vecs):  # the C ABI has no length arguments: a short buffer would be read past its end
            if n and v.size < (n - 1) * stride + 32:
                raise ValueError("scalar vector %d holds %d bytes, %d points at stride %d need %d" % (j, v.size, n, stride, (n - 1) * stride + 32))
        ptrs = (ctypes.c_void_p * max(k, 1))(*[v.ctypes.data for v in vecs])
        out = np.zeros((k, 72), dtype=np.uint8)
        _check(lib().cozk_msm_batch(self.handle, srs, base_offset, n, ptrs, k, stride, form, max_num_bits, _ptr(out)))
        return out

    def msm_batch_ptrs(self, srs, ptr_list, n, base_offset=0, stride=32, form=MONT, max_num_bits=0, device=None, out=None):
        """Raw-pointer variant: host pointers (device=None) or device pointers on context device `device`."""
        k = len(ptr_list)
        ptrs = (ctypes.c_void_p * max(k, 1))(*ptr_list)
        out = np.zeros((k, 72), dtype=np.uint8) if out is None else out
        if device is None:
            _check(lib().cozk_msm_batch(self.handle, srs, base_offset, n, ptrs, k, stride, form, max_num_bits, _ptr(out)))
        else:
            _check(lib().cozk_msm_batch_device(self.handle, device, srs, base_offset, n, ptrs, k, stride, form,
                                               max_num_bits, _ptr(out)))
        return out

    def msm_ragged(self, srs, ptr_list, offsets, lens, stride=32, form=MONT, device=0):
        """cozk_msm_ragged_device: vector j = lens[j] device-resident scalars at ptr_list[j] against bases offsets[j] .. of the SRS."""
        k = len(ptr_list)
        ptrs = (ctypes.c_void_p * k)(*ptr_list)
        offs = (ctypes.c_size_t * k)(*offsets)
        ln = (ctypes.c_size_t * k)(*lens)
        out = np.zeros((k, 72), dtype=np.uint8)
        _check(lib().cozk_msm_ragged_device(self.handle, device, srs, offs, ln, ptrs, k, stride, form, _ptr(out)))
        return out

    def fixed_base_batch_mul(self, base72, scalars, stride=32, form=MONT, register=False):
        """out[i] = scalars[i] * base (SRS generation, G1 half).  Returns (points (n, 72), srs handle or None)."""
        base72 = np.ascontiguousarray(base72, dtype=np.uint8).reshape(72)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
        n = scalars.size // stride
        out = np.zeros((n, 72), dtype=np.uint8)
        h = ctypes.c_uint64()
        _check(lib().cozk_fixed_base_batch_mul(self.handle, _ptr(base72), _ptr(scalars), n, stride, form, _ptr(out),
                                               ctypes.byref(h) if register else None))
        return out, (h.value if register else None)

    def set_option(self, name, value):
        _check(lib().cozk_set_option(self.handle, name.encode(), int(value)))

    def last_stats(self, device=0):
        s = (ctypes.c_double * 12)()
        _check(lib().cozk_last_stats_device(self.handle, device, s))
        keys = ["h2d_ms", "decompose_ms", "sort_ms", "accumulate_ms", "reduce_ms", "finish_ms", "total_ms", "launches",
                "window", "windows", "field_mults", "pairs"]
        return dict(zip(keys, list(s)))

    # ---- memory / generators / test kernels
    def alloc(self, nbytes, device=0):
        return DeviceBuffer(self, nbytes, device)

    def flush_l2(self, device=0):
        _check(lib().cozk_dev_flush_l2(self.handle, device))

    def testgen_bases(self, seed, n, start=0, device=0):
        buf = self.alloc(max(n, 1) * 64, device)
        _check(testlib().cozk_testgen_bases(self.handle, device, seed, start, n, ctypes.c_void_p(buf.ptr)))
        return buf

    def testgen_scalars(self, dist, seed, n, form=MONT, stride=32, start=0, total_n=None, device=0):
        buf = self.alloc(max(n, 1) * stride, device)
        _check(testlib().cozk_testgen_scalars(self.handle, device, DIST[dist], seed, start, n,
                                          n if total_n is None else total_n, form, ctypes.c_void_p(buf.ptr), stride))
        return buf

    def field_op(self, op, a, b=None, device=0):
        ops = {"mul": 0, "add": 1, "sub": 2, "sqr": 3, "inv": 4, "fr_from_mont": 5, "mul2_xyyx": 7}
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
        n = a.shape[0]
        da = self.alloc(a.nbytes, device).upload(a)
        db = self.alloc(a.nbytes, device).upload(np.ascontiguousarray(b, dtype=np.uint8)) if b is not None else None
        do = self.alloc(a.nbytes, device)
        try:
            _check(testlib().cozk_test_field_op(self.handle, device, ops[op], ctypes.c_void_p(da.ptr),
                                            ctypes.c_void_p(db.ptr) if db else None, ctypes.c_void_p(do.ptr), n))
            return do.download().reshape(n, 32)
        finally:
            for x in (da, db, do):
                if x:
                    x.free()

    def g1_op(self, op, a, b=None, device=0):
        ops = {"add": 0, "madd": 1, "dbl": 2}
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 72)
        n = a.shape[0]
        da = self.alloc(a.nbytes, device).upload(a)
        db = self.alloc(a.nbytes, device).upload(np.ascontiguousarray(b, dtype=np.uint8)) if b is not None else None
        do = self.alloc(a.nbytes, device)
        try:
            _check(testlib().cozk_test_g1_op(self.handle, device, ops[op], ctypes.c_void_p(da.ptr),
                                         ctypes.c_void_p(db.ptr) if db else None, ctypes.c_void_p(do.ptr), n))
            return do.download().reshape(n, 72)
        finally:
            for x in (da, db, do):
                if x:
                    x.free()

    def sort_pairs(self, keys, vals, key_bits, device=0):
        """The engine's pair sort on its own (include/cozk_test.h): (key, val) uint32 pairs grouped by key, keys ascending."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        if keys.size and int(keys.max()) >> key_bits:
            raise ValueError("keys must be below 2^key_bits")
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        m = keys.size
        bufs = [self.alloc(max(m, 1) * 4, device) for _ in range(4)]
        try:
            bufs[0].upload(keys.view(np.uint8))
            bufs[1].upload(vals.view(np.uint8))
            _check(testlib().cozk_test_sort(self.handle, device, ctypes.c_void_p(bufs[0].ptr), ctypes.c_void_p(bufs[1].ptr), m, key_bits,
                                        None, 0, 0, 0, 0, 0, 0, 0, 0, 0, ctypes.c_void_p(bufs[2].ptr), ctypes.c_void_p(bufs[3].ptr)))
            return bufs[2].download(m * 4).view(np.uint32), bufs[3].download(m * 4).view(np.uint32)
        finally:
            for b in bufs:
                b.free()

    def decompose_sort(self, dscalars, n, c, g=1, stride=32, form=MONT, windows=None, table_stride=0, val_offset=0, key_bits=0,
                       fused=True, device=0):
        """(keys, vals) of the plain decompose layout of g vectors of n device-resident scalars, sorted by key_bits key bits:
        fused=True through the engine's fused first pass, fused=False decompose kernel + generic passes (key_bits=0: unsorted)."""
        W = (254 + 1 + c - 1) // c if windows is None else windows
        m = g * n * W
        ko, vo = self.alloc(max(m, 1) * 4, device), self.alloc(max(m, 1) * 4, device)
        try:
            _check(testlib().cozk_test_sort(self.handle, device, None, None, 0, key_bits, ctypes.c_void_p(dscalars.ptr), n, g, stride, form,
                                        c, W, table_stride, val_offset, 1 if fused else 0, ctypes.c_void_p(ko.ptr), ctypes.c_void_p(vo.ptr)))
            return ko.download(m * 4).view(np.uint32), vo.download(m * 4).view(np.uint32)
        finally:
            ko.free()
            vo.free()

    def affine_rounds(self, keys, vals, dpts, total_buckets, rounds, reference=False, device=0, download=True):
        """The batched-affine pre-reduction on its own (include/cozk_test.h): `rounds` halving rounds over pairs grouped by key
        whose vals index into the device buffer dpts (64-byte affine points).  Returns (keys, vals, points[m, 64]) of the reduced
        list, the overflow lists [(keys, points)] per round, and the time of the rounds in ms."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        m = keys.size
        mo = m
        for _ in range(rounds):
            mo = (mo + 1) // 2
        cap = min((m + 1) // 2, total_buckets + 1)
        dk, dv = self.alloc(m * 4, device), self.alloc(m * 4, device)
        ko, vo, po = self.alloc(mo * 4, device), self.alloc(mo * 4, device), self.alloc(mo * 64, device)
        ok, op = (self.alloc(rounds * cap * 4, device), self.alloc(rounds * cap * 64, device)) if download else (None, None)
        try:
            dk.upload(keys.view(np.uint8))
            dv.upload(vals.view(np.uint8))
            out_m, ms = ctypes.c_size_t(), ctypes.c_double()
            counts = (ctypes.c_uint * rounds)()
            _check(testlib().cozk_test_affine_rounds(self.handle, device, ctypes.c_void_p(dk.ptr), ctypes.c_void_p(dv.ptr), m,
                                                     ctypes.c_void_p(dpts.ptr), total_buckets, rounds, 1 if reference else 0,
                                                     ctypes.c_void_p(ko.ptr), ctypes.c_void_p(vo.ptr), ctypes.c_void_p(po.ptr),
                                                     ctypes.byref(out_m), ctypes.c_void_p(ok.ptr) if ok else None,
                                                     ctypes.c_void_p(op.ptr) if op else None, cap, counts, ctypes.byref(ms)))
            assert out_m.value == mo
            if not download:
                return None, None, ms.value
            rk, rv = ko.download(mo * 4).view(np.uint32), vo.download(mo * 4).view(np.uint32)
            rp = po.download(mo * 64).reshape(mo, 64)
            allk, allp = ok.download(rounds * cap * 4).view(np.uint32), op.download(rounds * cap * 64).reshape(rounds * cap, 64)
            ovf = [(allk[r * cap:r * cap + counts[r]].copy(), allp[r * cap:r * cap + counts[r]].copy()) for r in range(rounds)]
            return (rk, rv, rp), ovf, ms.value
        finally:
            for b in (dk, dv, ko, vo, po, ok, op):
                if b:
                    b.free()

    def microbench(self, which, blocks, threads, iters, device=0):
        names = {"imad": 0, "fq_mul": 1, "fq_sqr": 2, "madd": 3, "imad_cc": 4, "imad_lo": 5, "imad_hi": 6, "fq_mul4": 7}
        ms, ops = ctypes.c_double(), ctypes.c_double()
        _check(testlib().cozk_microbench(self.handle, device, names[which], blocks, threads, iters, ctypes.byref(ms), ctypes.byref(ops)))
        return ms.value, ops.value


def g1_sum(points72):
    """Host-side sum of 72-byte wire points (combine_comm / combine_commitment_shares semantics)."""
    pts = np.ascontiguousarray(points72, dtype=np.uint8).reshape(-1, 72)
    out = np.zeros(72, dtype=np.uint8)
    _check(lib().cozk_g1_sum(_ptr(pts), pts.shape[0], _ptr(out)))
    return out


from . import spartan  # noqa: E402,F401  (co-spartan's worker-side commitment functions over the same entry points)
from . import pst13  # noqa: E402  (host-side mirror of the reference's PST13 / MultilinearPC interface)
from . import rep3  # noqa: E402  (device-resident Rep3 polynomials: include/cozk_rep3.h)
