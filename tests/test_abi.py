"""CPU tier: the C-ABI shared library loads and exports every symbol the headers declare; no compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(path):
    with open(path) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cozk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cozk):
    L = cozk.lib()
    declared = (_declared(os.path.join(ROOT, "include", "cozk_msm.h")) + _declared(os.path.join(ROOT, "include", "cozk_rep3.h"))
                + _declared(os.path.join(ROOT, "include", "cozk_pst13.h")))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(set(declared)) == sorted(cozk.ABI_SYMBOLS)


def test_test_entry_points_live_in_their_own_library(cozk):
    """include/cozk_test.h (generators, test kernels, microbenchmarks) is served by libcozk_test.so; the product library
    exports none of it."""
    T, L = cozk.testlib(), cozk.lib()
    declared = [n for n in _declared(os.path.join(ROOT, "include", "cozk_test.h"))]
    assert sorted(declared) == sorted(cozk.TEST_ABI_SYMBOLS)
    for name in declared:
        assert hasattr(T, name), name
        assert not hasattr(L, name), name


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "co-zkvms_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", ".inc")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/bn254" not in src, fn
                assert "liboracle" not in src, fn


def test_no_silent_fallback_without_gpu(cozk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cozk.CozkError) as e:
        cozk.Context()
    assert e.value.code == cozk.ERR_NO_DEVICE


def test_host_side_point_sum(cozk, orc):
    """cozk_g1_sum is pure host code (it combines per-GPU partial sums): check it against the oracle."""
    pts = np.zeros((9, 72), np.uint8)
    pts[:, :64] = orc.gen_bases(4, 9)
    acc = pts[0]
    for i in range(1, 9):
        acc = orc.g1_op("add", acc[None, :], pts[i][None, :])[0]
    assert (cozk.g1_sum(pts) == acc).all()
    # P + P, P + (-P), identities
    assert (cozk.g1_sum(np.stack([pts[0], pts[0]])) == orc.g1_op("dbl", pts[0][None, :])[0]).all()
    neg = orc.g1_op("neg", pts[0][None, :])[0]
    assert cozk.g1_sum(np.stack([pts[0], neg]))[64] == 1
    ident = H.point_wire(None)
    assert (cozk.g1_sum(np.stack([ident, pts[3], ident])) == pts[3]).all()
    assert cozk.g1_sum(np.zeros((0, 72), np.uint8))[64] == 1


def test_combine_commitment_shares_host(cozk, orc):
    pst = cozk.pst13
    pts = np.zeros((3, 72), np.uint8)
    pts[:, :64] = orc.gen_bases(6, 3)
    cs = [pst.PST13Commitment(10, pts[i]) for i in range(3)]
    got = pst.combine_commitment_shares(cs)
    want = orc.g1_op("add", orc.g1_op("add", pts[0][None], pts[1][None]), pts[2][None])[0]
    assert got.nv == 10 and (got.g_product == want).all()
    with pytest.raises(cozk.CozkError):
        pst.combine_commitment_shares([cs[0], pst.PST13Commitment(11, pts[1])])


def test_coordinate_prove_and_combine_comm_host(cozk, orc):
    pst = cozk.pst13
    pts = np.zeros((12, 72), np.uint8)
    pts[:, :64] = orc.gen_bases(8, 12)
    parties = [pts[0:4], pts[4:8], pts[8:12]]
    got = pst.coordinate_prove(parties)
    for i in range(4):
        want = orc.g1_op("add", orc.g1_op("add", pts[i][None], pts[4 + i][None]), pts[8 + i][None])[0]
        assert (got[i] == want).all()
    chunks = [pst.PST13Commitment(5, pts[i]) for i in range(4)]
    c = pst.combine_comm(chunks)
    want = pts[0]
    for i in range(1, 4):
        want = orc.g1_op("add", want[None], pts[i][None])[0]
    assert c.nv == 7 and (c.g_product == want).all()
    with pytest.raises(cozk.CozkError):
        pst.combine_comm(chunks[:3])
