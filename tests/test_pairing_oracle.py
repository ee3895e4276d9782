"""CPU tier: the pairing oracle (oracle/pairing.py) and the PST13 verifier equation it evaluates
(co-jolt/src/poly/commitment/pst13.rs:536-545: prove -> verify -> MultilinearPC::check).  The open() restated here uses
nothing but Python-integer curve arithmetic, so the equation is checked against an implementation that shares no code
with the engine; tests/test_gpu_pst13.py then holds the engine's proofs to the same equation."""
import pytest

from oracle import pairing as pr
from oracle import pyref

R = pr.R


def test_pairing_is_bilinear_and_non_degenerate():
    assert pr.g2_is_on_curve(pr.G2)
    assert pr.g2_mul(pr.G2, R - 1) == pr.g2_neg(pr.G2)          # r * G2 = O
    e1 = pr.pairing(pyref.G1, pr.G2)
    assert e1 != pr.F12_ONE and pr.f12_pow(e1, R) == pr.F12_ONE  # an element of order r
    a, b = 0x1F2E3D4C5B6A7988, 0x123456789ABCDEF0123456789
    assert pr.pairing(pyref.mul(a, pyref.G1), pr.G2) == pr.f12_pow(e1, a)
    assert pr.pairing(pyref.G1, pr.g2_mul(pr.G2, b)) == pr.f12_pow(e1, b)
    assert pr.pairing(pyref.mul(a, pyref.G1), pr.g2_mul(pr.G2, b)) == pr.f12_pow(e1, a * b % R)
    p2 = pyref.base_point(3, 7)
    lhs = pr.pairing(pyref.add(pyref.G1, p2), pr.G2)
    assert lhs == pr.f12_mul(e1, pr.pairing(p2, pr.G2))
    assert pr.pairing(None, pr.G2) == pr.F12_ONE and pr.pairing(pyref.G1, None) == pr.F12_ONE
    assert pr.pairing_product_is_one([(pyref.G1, pr.G2), (pyref.neg(pyref.G1), pr.G2)])


def open_python(levels, evals, point):
    """open() (pst13.rs:428-474) on Python integers and affine chord-and-tangent arithmetic."""
    nv = len(point)
    r = list(evals)
    proofs = []
    for i in range(nv):
        half = 1 << (nv - i - 1)
        q = [(r[2 * b + 1] - r[2 * b]) % R for b in range(half)]
        r = [(r[2 * b] * (1 - point[i]) + r[2 * b + 1] * point[i]) % R for b in range(half)]
        proofs.append(pyref.msm_naive([q[x >> 1] for x in range(2 * half)], levels[i]))
    return proofs, r[0]


@pytest.mark.parametrize("nv", [1, 2, 3])
def test_verifier_equation_accepts_open_and_rejects_tampering(nv):
    levels, vk, t = pr.setup(nv, seed=40 + nv)
    evals = [pyref.scalar_uniform(50 + nv, i) for i in range(1 << nv)]
    point = [pyref.scalar_uniform(60 + nv, i) for i in range(nv)]
    commitment = pyref.msm_naive(evals, levels[0])
    # the SRS is the eq basis of the trapdoor: the commitment is f(t) * g
    f_t = 0
    for x, e in enumerate(evals):
        w = e
        for j in range(nv):
            w = w * (t[j] if (x >> j) & 1 else (1 - t[j])) % R
        f_t = (f_t + w) % R
    assert commitment == pyref.mul(f_t, pyref.G1)
    proofs, value = open_python(levels, evals, point)
    assert pr.verify_opening(vk, commitment, point, value, proofs)
    assert not pr.verify_opening(vk, commitment, point, (value + 1) % R, proofs)
    if nv > 1:
        assert not pr.verify_opening(vk, commitment, point[::-1], value, proofs)          # point order matters
        assert not pr.verify_opening(vk, commitment, point, value, proofs[::-1])          # level order matters
    bad = list(proofs)
    bad[0] = pyref.add(bad[0], pyref.G1)
    assert not pr.verify_opening(vk, commitment, point, value, bad)
