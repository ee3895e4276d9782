"""CPU tier: the generated inline-PTX field arithmetic, interpreted instruction by instruction in Python and
compared with big-integer arithmetic; and the generated file is what the generator emits now."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_field_ptx", os.path.join(ROOT, "tools", "gen_field_ptx.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_ptx_programs_against_bigint():
    assert _gen().self_check(trials=400, seed=7)


def test_generated_file_is_current(tmp_path):
    g = _gen()
    out = tmp_path / "field_ptx.inc"
    import sys
    argv = sys.argv
    sys.argv = ["gen", "-o", str(out)]
    try:
        g.main()
    finally:
        sys.argv = argv
    with open(os.path.join(ROOT, "co-zkvms_b200", "csrc", "field_ptx.inc")) as f:
        assert f.read() == out.read_text()


def test_multiply_count():
    g = _gen()
    pg = g.gen_mul(g.P)
    wide = sum(1 for i in pg.ins if i[0] == "mul.wide")
    pairs = sum(1 for i in pg.ins if i[0].startswith("mad") and ".lo" in i[0])
    single = sum(1 for i in pg.ins if i[0] == "mul.lo")
    # 8 rows x (8 a*b + 8 m*p) products + 8 m = 136 integer multiply-adds per field multiplication
    assert wide + pairs + single == 136


def test_squaring_multiply_count():
    """Squaring: 28 cross products + 8 squares + 8 rows x (8 m*p + 1 m) = 108 multiply-adds (a multiplication: 136)."""
    g = _gen()
    pg = g.gen_sqr_sos(g.P)
    pairs = sum(1 for i in pg.ins if i[0].startswith("mad") and ".lo" in i[0])
    single = sum(1 for i in pg.ins if i[0] == "mul.lo")
    assert pairs + single == 108


def test_fused_product_pair_multiply_count():
    """a*b + c*d with one reduction: 8 rows x (8 + 8 + 8 m*p + 1 m) = 200 multiply-adds (two multiplications: 272)."""
    g = _gen()
    pg = g.gen_mul2(g.P)
    wide = sum(1 for i in pg.ins if i[0] == "mul.wide")
    pairs = sum(1 for i in pg.ins if i[0].startswith("mad") and ".lo" in i[0])
    single = sum(1 for i in pg.ins if i[0] == "mul.lo")
    assert wide + pairs + single == 200
