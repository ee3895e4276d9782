"""GPU tier: the PTX field arithmetic and the XYZZ group law on the device, bit-exact against the oracle."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _rand_field(n, mod, seed):
    rng = np.random.default_rng(seed)
    edge = [0, 1, 2, mod - 1, mod - 2, 1 << 253, (1 << 254) % mod, mod >> 1, 0xFFFFFFFF, (1 << 224) - 1]
    vals = edge + [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(n - len(edge))]
    return np.stack([H.le32(v) for v in vals])


def test_fq_ops_2pow16(ctx, orc):
    n = 1 << 16
    a = _rand_field(n, H.P, 1)
    b = _rand_field(n, H.P, 2)[::-1].copy()
    for op in ("mul", "add", "sub"):
        want = orc.field_op("fq", op, a.view(np.uint64), b.view(np.uint64)).view(np.uint8)
        assert (ctx.field_op(op, a, b) == want).all(), op
    assert (ctx.field_op("sqr", a) == orc.field_op("fq", "sqr", a.view(np.uint64)).view(np.uint8)).all()
    # the fused product pair (x*y + y*x with one reduction) = 2xy
    prod = orc.field_op("fq", "mul", a.view(np.uint64), b.view(np.uint64))
    assert (ctx.field_op("mul2_xyyx", a, b) == orc.field_op("fq", "add", prod, prod).view(np.uint8)).all()


def test_fq_inv_and_fr_from_mont(ctx, orc):
    a = _rand_field(2048, H.P, 3)
    assert (ctx.field_op("inv", a) == orc.field_op("fq", "inv", a.view(np.uint64)).view(np.uint8)).all()
    s = _rand_field(4096, H.R, 4)
    assert (ctx.field_op("fr_from_mont", s) == orc.field_op("fr", "from_mont", s.view(np.uint64)).view(np.uint8)).all()


def test_group_law_golden_and_random(ctx, orc):
    for case in H.golden()["adds"]:
        a, b = H.point_wire(H.parse_point(case["a"])), H.point_wire(H.parse_point(case["b"]))
        want = H.point_wire(H.parse_point(case["sum"]))
        assert (ctx.g1_op("add", a, b)[0] == want).all(), case
        if case["b"] is not None:
            assert (ctx.g1_op("madd", a, b)[0] == want).all(), case
    n = 4096
    pts = np.zeros((n, 72), np.uint8)
    pts[:, :64] = orc.gen_bases(2, n)
    other = np.roll(pts, 1, axis=0)
    other[::7] = pts[::7]                                   # P + P
    other[3::11] = orc.g1_op("neg", pts[3::11])             # P + (-P)
    other[5::13] = H.point_wire(None)                       # P + O
    assert (ctx.g1_op("add", pts, other) == orc.g1_op("add", pts, other)).all()
    fin = other.copy()
    fin[5::13] = pts[6::13][: fin[5::13].shape[0]]
    assert (ctx.g1_op("madd", pts, fin) == orc.g1_op("add", pts, fin)).all()
    assert (ctx.g1_op("dbl", pts) == orc.g1_op("dbl", pts)).all()


def test_device_generators_match_oracle(ctx, orc):
    n = 3000
    d = ctx.testgen_bases(1, n)
    assert (d.download().reshape(n, 64) == orc.gen_bases(1, n)).all()
    d.free()
    d = ctx.testgen_bases(9, 100, start=12345)
    assert (d.download().reshape(100, 64) == orc.gen_bases(9, 100, start=12345)).all()
    d.free()
    for dist in ("uniform", "const", "wminus", "dup", "small16", "zero_half"):
        for form in (0, 1):
            d = ctx.testgen_scalars(dist, 5, n, form=form)
            assert (d.download().reshape(n, 32) == orc.gen_scalars(dist, 5, n, form=form)).all(), (dist, form)
            d.free()
    d = ctx.testgen_scalars("uniform", 5, 500, stride=64, start=77, total_n=1000)
    got = d.download().reshape(500, 64)[:, :32]
    assert (got == orc.gen_scalars("uniform", 5, 500, start=77, total_n=1000)).all()
    d.free()
