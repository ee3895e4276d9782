#!/usr/bin/env python3
"""Generator (and checker) for the 8x32-bit-limb Montgomery field arithmetic used by the CUDA kernels.

It emits co-zkvms_b200/csrc/field_ptx.inc: one inline-PTX block per operation (mul, sqr, add, sub for
BN254 Fq; mul for Fr), so that the carry flag never lives across two asm statements.  The same instruction
lists can be *interpreted* here in Python (a tiny PTX subset: mul/mad/madc/add/addc/sub/subc/selp on u32
with one carry flag), which is how the generated code is verified against big-integer arithmetic without a
GPU (tests/test_field_ptx.py).

Montgomery product, operand scanning, one limb of b per row; the running total T is kept as TWO multi-limb
numbers so that every 32x32->64 product lands on an aligned (lo, hi) register pair and each row is a plain
carry chain of wide multiply-adds:
    T = EV + OD * 2^32,   EV = sum ev[k] 2^(32k),   OD = sum od[k] 2^(32k)
    row i:  EV += sum_{j even} a_j b_i 2^(32 j)          (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32[.X])
            OD += sum_{j odd}  a_j b_i 2^(32 (j-1))
            m   = ev[0] * n0 mod 2^32
            EV += sum_{j even} m p_j 2^(32 j);  OD += sum_{j odd} m p_j 2^(32 (j-1))     -> ev[0] == 0
            T  /= 2^32:   EV' = OD,  OD' = EV >> 64,  and the stray limb ev[1] (weight 2^0) is folded into
                          ev'[0] at the start of the next row, its carry feeding the OD' chain (weight 2^32).
ptxas fuses each (mad.lo.cc, madc.hi.cc) pair into one IMAD.WIDE.U32 with a carry predicate, so a product costs
8 rows x (8 + 8 + 1) = 136 integer multiply-adds.  T < 2^288 throughout (254-bit modulus), hence ev needs 9
limbs, od 8, and no chain ever carries out of its top limb.
"""
import argparse
import os
import random

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
M32 = 0xFFFFFFFF


def limbs32(x, n=8):
    return [(x >> (32 * k)) & M32 for k in range(n)]


class Prog:
    """A straight-line PTX program over named u32 registers."""

    def __init__(self):
        self.ins = []
        self.tmp = 0
        self.allow_wrap = False  # True: a chain top may carry out (modular subtraction adds the modulus back mod 2^256)

    def new(self, prefix="t"):
        self.tmp += 1
        return "%s%d" % (prefix, self.tmp)

    def emit(self, op, dst, *src):
        self.ins.append((op, dst) + tuple(src))

    # ---- interpreter
    def run(self, env):
        env = dict(env)
        cf = 0

        def val(x):
            return x if isinstance(x, int) else env[x]

        for ins in self.ins:
            op, dst, src = ins[0], ins[1], [val(s) for s in ins[2:]]
            if op == "mul.wide":  # dst = (lo, hi)
                prod = src[0] * src[1]
                env[dst[0]], env[dst[1]] = prod & M32, prod >> 32
                continue
            if op == "mov":
                r = src[0]
            elif op == "mul.lo":
                r = (src[0] * src[1]) & M32
            elif op == "mul.hi":
                r = (src[0] * src[1]) >> 32
            elif op in ("mad.lo", "mad.lo.cc", "madc.lo", "madc.lo.cc", "mad.hi", "mad.hi.cc", "madc.hi", "madc.hi.cc"):
                prod = src[0] * src[1]
                part = (prod & M32) if ".lo" in op else (prod >> 32)
                s = part + src[2] + (cf if op.startswith("madc") else 0)
                r = s & M32
                if op.endswith(".cc"):
                    cf = s >> 32
                elif op.startswith("madc"):
                    assert self.allow_wrap or s >> 32 == 0, "carry out of a chain top"
            elif op in ("add", "add.cc", "addc", "addc.cc"):
                s = src[0] + src[1] + (cf if op.startswith("addc") else 0)
                r = s & M32
                if op.endswith(".cc"):
                    cf = s >> 32
                elif op == "addc":
                    assert self.allow_wrap or s >> 32 == 0, "carry out of a chain top"
            elif op in ("sub", "sub.cc", "subc", "subc.cc"):
                s = src[0] - src[1] - (cf if op.startswith("subc") else 0)
                r = s & M32
                if op.endswith(".cc"):
                    cf = 1 if s < 0 else 0
            elif op == "selp.ne0":  # dst = (src[2] != 0) ? src[0] : src[1]
                r = src[0] if src[2] != 0 else src[1]
            elif op == "and":
                r = src[0] & src[1]
            elif op == "xor":
                r = src[0] ^ src[1]
            elif op == "addc.wrap":  # top limb of a two's-complement sum: the carry out is meant to be dropped
                r = (src[0] + src[1] + cf) & M32
            else:
                raise ValueError(op)
            env[dst] = r
        return env

    # ---- PTX text
    def ptx(self, names):
        """names: register -> operand placeholder ('%0'...) for inputs/outputs; everything else becomes a local .reg."""
        local = []
        seen = set()
        for ins in self.ins:
            flat = []
            for x in ins[1:]:
                flat.extend(x if isinstance(x, tuple) else [x])
            for x in flat:
                if isinstance(x, str) and x not in names and x not in seen:
                    seen.add(x)
                    local.append(x)
        lines = ["{"]
        for i in range(0, len(local), 8):
            lines.append(".reg .u32 " + ", ".join(local[i:i + 8]) + ";")
        if any(ins[0] == "mul.wide" for ins in self.ins):
            lines.append(".reg .u64 wd;")
        need_pred = any(ins[0] == "selp.ne0" for ins in self.ins)
        if need_pred:
            lines.append(".reg .pred pq;")
        pred_set_for = None

        def o(x):
            if isinstance(x, int):
                return "0x%08x" % x
            return names.get(x, x)

        for ins in self.ins:
            op, dst, src = ins[0], ins[1], ins[2:]
            if op == "mul.wide":
                lines.append("mul.wide.u32 wd, %s, %s;" % (o(src[0]), o(src[1])))
                lines.append("mov.b64 {%s, %s}, wd;" % (o(dst[0]), o(dst[1])))
            elif op == "selp.ne0":
                if pred_set_for != src[2]:
                    lines.append("setp.ne.u32 pq, %s, 0;" % o(src[2]))
                    pred_set_for = src[2]
                lines.append("selp.u32 %s, %s, %s, pq;" % (o(dst), o(src[0]), o(src[1])))
            elif op == "mov":
                lines.append("mov.u32 %s, %s;" % (o(dst), o(src[0])))
            elif op == "and":
                lines.append("and.b32 %s, %s, %s;" % (o(dst), o(src[0]), o(src[1])))
            elif op == "xor":
                lines.append("xor.b32 %s, %s, %s;" % (o(dst), o(src[0]), o(src[1])))
            elif op == "addc.wrap":
                lines.append("addc.u32 %s, %s, %s;" % (o(dst), o(src[0]), o(src[1])))
            else:
                lines.append("%s.u32 %s, %s;" % (op, o(dst), ", ".join(o(s) for s in src)))
        lines.append("}")
        return lines


def cond_sub(pg, t, mod, out):
    """out = t - mod if t >= mod else t   (t: 8 register names)"""
    ml = limbs32(mod)
    s = [pg.new("s") for _ in range(8)]
    for k in range(8):
        pg.emit("sub.cc" if k == 0 else "subc.cc", s[k], t[k], ml[k])
    brw = pg.new("w")
    pg.emit("subc", brw, 0, 0)  # 0xffffffff when t < mod
    for k in range(8):
        pg.emit("selp.ne0", out[k], t[k], s[k], brw)


def gen_mul(mod, square=False):
    """r = a*b/2^256 mod `mod`, fully reduced.  Inputs a0..a7, b0..b7 (b = a for square), outputs r0..r7."""
    pg = Prog()
    a = ["a%d" % k for k in range(8)]
    b = a if square else ["b%d" % k for k in range(8)]
    ml = limbs32(mod)
    n0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    ev = [pg.new("e") for _ in range(9)]
    od = [pg.new("o") for _ in range(8)]
    stray = None
    for i in range(8):
        bi = b[i]
        if i == 0:
            for k in range(4):
                pg.emit("mul.wide", (ev[2 * k], ev[2 * k + 1]), a[2 * k], bi)
                pg.emit("mul.wide", (od[2 * k], od[2 * k + 1]), a[2 * k + 1], bi)
            pg.emit("mov", ev[8], 0)
        else:
            # fold the stray limb (weight 2^0); its carry has weight 2^32 = od[0]
            pg.emit("add.cc", ev[0], ev[0], stray)
            for k in range(4):
                pg.emit("madc.lo.cc", od[2 * k], a[2 * k + 1], bi, od[2 * k])
                # the top pair never carries out (T < 2^288); keep .cc so that ptxas still fuses it into IMAD.WIDE
                pg.emit("madc.hi.cc", od[2 * k + 1], a[2 * k + 1], bi, od[2 * k + 1])
            for k in range(4):
                pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", ev[2 * k], a[2 * k], bi, ev[2 * k])
                pg.emit("madc.hi.cc", ev[2 * k + 1], a[2 * k], bi, ev[2 * k + 1])
            pg.emit("addc", ev[8], ev[8], 0)
        m = pg.new("m")
        pg.emit("mul.lo", m, ev[0], n0)
        for k in range(4):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", ev[2 * k], m, ml[2 * k], ev[2 * k])
            pg.emit("madc.hi.cc", ev[2 * k + 1], m, ml[2 * k], ev[2 * k + 1])
        pg.emit("addc", ev[8], ev[8], 0)
        for k in range(4):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", od[2 * k], m, ml[2 * k + 1], od[2 * k])
            pg.emit("madc.hi.cc", od[2 * k + 1], m, ml[2 * k + 1], od[2 * k + 1])
        # T /= 2^32: ev[0] is zero now; ev[1] becomes the stray limb; EV' = OD; OD' = ev[2..8] ++ [0]
        stray = ev[1]
        new_od = ev[2:9]
        z = pg.new("z")
        pg.emit("mov", z, 0)
        new_od = new_od + [z]
        z8 = pg.new("z")
        pg.emit("mov", z8, 0)
        ev, od = od + [z8], new_od
    # result = stray + EV + OD*2^32  as one 8-limb addition: (stray, od[0..6]) + ev[0..7]; od[7] and ev[8] are zero
    lo = [stray] + od[0:7]
    t = [pg.new("r") for _ in range(8)]
    for k in range(8):
        pg.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), t[k], ev[k], lo[k])
    cond_sub(pg, t, mod, ["r%d" % k for k in range(8)])
    return pg


def gen_mul2(mod):
    """r = (a*b + c*d) / 2^256 mod `mod`, fully reduced: both products share ONE Montgomery reduction, so the pair costs
    8 rows x (8 + 8 + 8 + 1) = 200 multiply-adds instead of 272 (the Y coordinate of every group addition is such a
    sum: R*(Q - X3) + (-Y1)*PPP).  Same EV/OD split as gen_mul; row i adds a*b_i and c*d_i before the reduction row.
    T < (3 * 2^32 + 3) * mod < 2^288 throughout and the result is < 2*mod^2/2^256 + mod < 2*mod: one conditional
    subtraction."""
    pg = Prog()
    a = ["a%d" % k for k in range(8)]
    b = ["b%d" % k for k in range(8)]
    c = ["c%d" % k for k in range(8)]
    d = ["d%d" % k for k in range(8)]
    ml = limbs32(mod)
    n0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    ev = [pg.new("e") for _ in range(9)]
    od = [pg.new("o") for _ in range(8)]
    stray = None
    for i in range(8):
        bi, di = b[i], d[i]
        if i == 0:
            for k in range(4):
                pg.emit("mul.wide", (ev[2 * k], ev[2 * k + 1]), a[2 * k], bi)
                pg.emit("mul.wide", (od[2 * k], od[2 * k + 1]), a[2 * k + 1], bi)
            pg.emit("mov", ev[8], 0)
        else:
            pg.emit("add.cc", ev[0], ev[0], stray)
            for k in range(4):
                pg.emit("madc.lo.cc", od[2 * k], a[2 * k + 1], bi, od[2 * k])
                pg.emit("madc.hi.cc", od[2 * k + 1], a[2 * k + 1], bi, od[2 * k + 1])
            for k in range(4):
                pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", ev[2 * k], a[2 * k], bi, ev[2 * k])
                pg.emit("madc.hi.cc", ev[2 * k + 1], a[2 * k], bi, ev[2 * k + 1])
            pg.emit("addc", ev[8], ev[8], 0)
        for k in range(4):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", od[2 * k], c[2 * k + 1], di, od[2 * k])
            pg.emit("madc.hi.cc", od[2 * k + 1], c[2 * k + 1], di, od[2 * k + 1])
        for k in range(4):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", ev[2 * k], c[2 * k], di, ev[2 * k])
            pg.emit("madc.hi.cc", ev[2 * k + 1], c[2 * k], di, ev[2 * k + 1])
        pg.emit("addc", ev[8], ev[8], 0)
        m = pg.new("m")
        pg.emit("mul.lo", m, ev[0], n0)
        for k in range(4):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", ev[2 * k], m, ml[2 * k], ev[2 * k])
            pg.emit("madc.hi.cc", ev[2 * k + 1], m, ml[2 * k], ev[2 * k + 1])
        pg.emit("addc", ev[8], ev[8], 0)
        for k in range(4):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", od[2 * k], m, ml[2 * k + 1], od[2 * k])
            pg.emit("madc.hi.cc", od[2 * k + 1], m, ml[2 * k + 1], od[2 * k + 1])
        stray = ev[1]
        new_od = ev[2:9]
        z = pg.new("z")
        pg.emit("mov", z, 0)
        new_od = new_od + [z]
        z8 = pg.new("z")
        pg.emit("mov", z8, 0)
        ev, od = od + [z8], new_od
    lo = [stray] + od[0:7]
    t = [pg.new("r") for _ in range(8)]
    for k in range(8):
        pg.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), t[k], ev[k], lo[k])
    cond_sub(pg, t, mod, ["r%d" % k for k in range(8)])
    return pg


def run_quadop(pg, x, y, z, w):
    env = {}
    for k in range(8):
        env["a%d" % k] = (x >> (32 * k)) & M32
        env["b%d" % k] = (y >> (32 * k)) & M32
        env["c%d" % k] = (z >> (32 * k)) & M32
        env["d%d" % k] = (w >> (32 * k)) & M32
    out = pg.run(env)
    return sum(out["r%d" % k] << (32 * k) for k in range(8))


def redc_tail(pg, t, mod, out):
    """out = (t[0..15] as a 512-bit integer) / 2^256 mod `mod`, fully reduced, for t < mod * 2^256 + small.
    t / R = redc(t_lo) + t_hi: the eight reduction rows of gen_mul without the multiplication rows (72 multiply-adds),
    then one 8-limb addition of the high half.  The stray limb of the previous row is folded into ev[0] first and its
    carry (weight 2^32) is picked up by the ODD chain of m * p, which therefore runs before the even one; mul.lo in
    between does not touch the carry flag."""
    ml = limbs32(mod)
    n0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    ev = [pg.new("e") for _ in range(9)]
    od = [pg.new("o") for _ in range(8)]
    for k in range(8):
        pg.emit("mov", ev[k], t[k])
        pg.emit("mov", od[k], 0)
    pg.emit("mov", ev[8], 0)
    stray = None
    for i in range(8):
        if stray is not None:
            pg.emit("add.cc", ev[0], ev[0], stray)
        m = pg.new("m")
        pg.emit("mul.lo", m, ev[0], n0)
        for k in range(4):
            first = "mad.lo.cc" if (k == 0 and stray is None) else "madc.lo.cc"
            pg.emit(first, od[2 * k], m, ml[2 * k + 1], od[2 * k])
            pg.emit("madc.hi.cc", od[2 * k + 1], m, ml[2 * k + 1], od[2 * k + 1])
        for k in range(4):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", ev[2 * k], m, ml[2 * k], ev[2 * k])
            pg.emit("madc.hi.cc", ev[2 * k + 1], m, ml[2 * k], ev[2 * k + 1])
        pg.emit("addc", ev[8], ev[8], 0)
        stray = ev[1]
        new_od = ev[2:9]
        z = pg.new("z")
        pg.emit("mov", z, 0)
        new_od = new_od + [z]
        z8 = pg.new("z")
        pg.emit("mov", z8, 0)
        ev, od = od + [z8], new_od
    lo = [stray] + od[0:7]
    u = [pg.new("u") for _ in range(8)]
    for k in range(8):
        pg.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), u[k], ev[k], lo[k])
    v = [pg.new("v") for _ in range(8)]
    for k in range(8):
        pg.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), v[k], u[k], t[8 + k])
    cond_sub(pg, v, mod, out)


def sqrwide(pg, a):
    """t[0..15] = a^2 as a plain integer with 36 multiply-adds: the 28 cross products a_i a_j (i < j) land on aligned
    pairs of two accumulators (EV: i + j even, OD: i + j odd, weight 2^32), their sum is doubled by one carry chain, and
    the eight squares a_i^2 are added by one chain over all sixteen limbs."""
    ev = [pg.new("e") for _ in range(16)]
    od = [pg.new("o") for _ in range(16)]
    for k in range(16):
        pg.emit("mov", ev[k], 0)
        pg.emit("mov", od[k], 0)

    def chain(acc, start, js, ai):
        for k, j in enumerate(js):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", acc[start + 2 * k], a[j], ai, acc[start + 2 * k])
            pg.emit("madc.hi.cc", acc[start + 2 * k + 1], a[j], ai, acc[start + 2 * k + 1])
        top = start + 2 * len(js)
        if js and top <= 15:
            pg.emit("addc", acc[top], acc[top], 0)

    for i in range(7):
        ev_js = [j for j in range(i + 1, 8) if (i + j) % 2 == 0]
        od_js = [j for j in range(i + 1, 8) if (i + j) % 2 == 1]
        if ev_js:
            chain(ev, i + ev_js[0], ev_js, a[i])          # positions i + j, contiguous pairs
        if od_js:
            chain(od, i + od_js[0] - 1, od_js, a[i])      # positions i + j -> OD index i + j - 1
    s = [pg.new("s") for _ in range(16)]
    pg.emit("mov", s[0], ev[0])
    for k in range(1, 16):
        pg.emit("add.cc" if k == 1 else ("addc.cc" if k < 15 else "addc"), s[k], ev[k], od[k - 1])
    d = [pg.new("d") for _ in range(16)]
    for k in range(16):
        pg.emit("add.cc" if k == 0 else ("addc.cc" if k < 15 else "addc"), d[k], s[k], s[k])
    for i in range(8):
        pg.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", d[2 * i], a[i], a[i], d[2 * i])
        pg.emit("madc.hi.cc" if i < 7 else "madc.hi", d[2 * i + 1], a[i], a[i], d[2 * i + 1])
    return d


def mulwide_n(pg, a, b):
    """len(a) == len(b) == n (even): returns 2n registers holding a * b as a plain integer, n^2 wide multiply-adds on the
    EV/OD split of gen_mulwide plus one merging addition."""
    n = len(a)
    ev = [pg.new("e") for _ in range(2 * n)]
    od = [pg.new("o") for _ in range(2 * n)]
    for k in range(n // 2):
        pg.emit("mul.wide", (ev[2 * k], ev[2 * k + 1]), a[2 * k], b[0])
        pg.emit("mul.wide", (od[2 * k], od[2 * k + 1]), a[2 * k + 1], b[0])
    for k in range(n, 2 * n):
        pg.emit("mov", ev[k], 0)
        pg.emit("mov", od[k], 0)

    def chain(acc, start, mults, bi):
        for k, aj in enumerate(mults):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", acc[start + 2 * k], aj, bi, acc[start + 2 * k])
            pg.emit("madc.hi.cc", acc[start + 2 * k + 1], aj, bi, acc[start + 2 * k + 1])
        if start + n <= 2 * n - 1:
            pg.emit("addc", acc[start + n], acc[start + n], 0)

    even_a, odd_a = a[0::2], a[1::2]
    for i in range(1, n):
        if i % 2 == 0:
            chain(ev, i, even_a, b[i])
            chain(od, i, odd_a, b[i])
        else:
            chain(ev, i + 1, odd_a, b[i])
            chain(od, i - 1, even_a, b[i])
    r = [ev[0]] + [pg.new("p") for _ in range(2 * n - 1)]
    for k in range(1, 2 * n):
        pg.emit("add.cc" if k == 1 else ("addc.cc" if k < 2 * n - 1 else "addc"), r[k], ev[k], od[k - 1])
    return r


def abs_diff(pg, x, y):
    """(|x - y| as len(x) registers, mask register: 0xffffffff when x < y else 0)"""
    n = len(x)
    d = [pg.new("d") for _ in range(n)]
    for k in range(n):
        pg.emit("sub.cc" if k == 0 else "subc.cc", d[k], x[k], y[k])
    m = pg.new("w")
    pg.emit("subc", m, 0, 0)
    out = [pg.new("d") for _ in range(n)]
    for k in range(n):
        pg.emit("xor", d[k], d[k], m)
    for k in range(n):  # (d ^ m) - m: adds one when m is all ones
        pg.emit("sub.cc" if k == 0 else ("subc.cc" if k < n - 1 else "subc"), out[k], d[k], m)
    return out, m


def karatsuba_wide(pg, a, b):
    """a * b (8 x 8 limbs -> 16) with three 4 x 4 products (48 wide multiply-adds instead of 64):
    a*b = z0 + (z0 + z2 + (a0 - a1)(b1 - b0)) 2^128 + z2 2^256,  z0 = a0 b0, z2 = a1 b1; the middle product is taken on
    absolute values and added or subtracted by sign (two's complement on 9 limbs)."""
    a0, a1, b0, b1 = a[0:4], a[4:8], b[0:4], b[4:8]
    z0 = mulwide_n(pg, a0, b0)
    z2 = mulwide_n(pg, a1, b1)
    da, sa = abs_diff(pg, a0, a1)
    db, sb = abs_diff(pg, b1, b0)
    z1 = mulwide_n(pg, da, db)
    s = pg.new("w")
    pg.emit("xor", s, sa, sb)  # all ones: the middle product is negative
    mid = [pg.new("m") for _ in range(9)]
    for k in range(8):
        pg.emit("add.cc" if k == 0 else "addc.cc", mid[k], z0[k], z2[k])
    pg.emit("addc", mid[8], 0, 0)
    x = [pg.new("x") for _ in range(8)]
    for k in range(8):
        pg.emit("xor", x[k], z1[k], s)
    cy = pg.new("x")
    pg.emit("add.cc", cy, s, 1)  # carry = 1 exactly when s is all ones: the +1 of the two's complement
    for k in range(8):
        pg.emit("addc.cc", mid[k], mid[k], x[k])
    pg.emit("addc.wrap", mid[8], mid[8], s)
    hi = z0[4:8] + z2[0:8]
    t = list(z0[0:4]) + [pg.new("t") for _ in range(12)]
    for k in range(12):
        op = "add.cc" if k == 0 else ("addc.cc" if k < 11 else "addc")
        pg.emit(op, t[4 + k], hi[k], mid[k] if k < 9 else 0)
    return t


def gen_mul_karatsuba(mod):
    """r = a*b/2^256 mod `mod`: Karatsuba product (48) + separate Montgomery reduction (72) = 120 multiply-adds.
    EXPERIMENT, not emitted: measured on B200 at 64.2 G products/s against 66.5 G/s for gen_mul (136 multiply-adds) -
    the ~90 extra additions and the longer dependency chains cost more than the 16 multiply-adds they save, and the
    mixed addition built on it spills at 128 registers (4.3 against 5.6 G additions/s)."""
    pg = Prog()
    a = ["a%d" % k for k in range(8)]
    b = ["b%d" % k for k in range(8)]
    t = karatsuba_wide(pg, a, b)
    redc_tail(pg, t, mod, ["r%d" % k for k in range(8)])
    return pg


def gen_sqr_sos(mod):
    """r = a^2 / 2^256 mod `mod`: 36 + 72 = 108 multiply-adds instead of 136."""
    pg = Prog()
    a = ["a%d" % k for k in range(8)]
    t = sqrwide(pg, a)
    redc_tail(pg, t, mod, ["r%d" % k for k in range(8)])
    return pg


def gen_mulwide():
    """r[0..15] = a * b as a plain 512-bit integer (no reduction): the product half of a lazily reduced
    multiply-accumulate (rep3_kernels.cuh: linear combinations of shared polynomials, chi dot products).  Same EV/OD
    split as gen_mul so that every product lands on an aligned register pair: EV collects the products whose limb
    position i+j is even, OD (weight 2^32) the odd ones.  64 wide multiply-adds."""
    pg = Prog()
    a = ["a%d" % k for k in range(8)]
    b = ["b%d" % k for k in range(8)]
    ev = [pg.new("e") for _ in range(16)]
    od = [pg.new("o") for _ in range(16)]
    for k in range(4):
        pg.emit("mul.wide", (ev[2 * k], ev[2 * k + 1]), a[2 * k], b[0])
        pg.emit("mul.wide", (od[2 * k], od[2 * k + 1]), a[2 * k + 1], b[0])
    for k in range(8, 16):
        pg.emit("mov", ev[k], 0)
        pg.emit("mov", od[k], 0)

    def chain(acc, start, mults, bi):
        for k, aj in enumerate(mults):
            pg.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", acc[start + 2 * k], aj, bi, acc[start + 2 * k])
            pg.emit("madc.hi.cc", acc[start + 2 * k + 1], aj, bi, acc[start + 2 * k + 1])
        if start + 8 <= 15:
            pg.emit("addc", acc[start + 8], acc[start + 8], 0)

    for i in range(1, 8):
        even_a, odd_a = [a[0], a[2], a[4], a[6]], [a[1], a[3], a[5], a[7]]
        if i % 2 == 0:
            chain(ev, i, even_a, b[i])      # positions i+2k (even)
            chain(od, i, odd_a, b[i])       # positions i+2k+1 -> OD index i+2k
        else:
            chain(ev, i + 1, odd_a, b[i])   # positions i+2k+1 (even)
            chain(od, i - 1, even_a, b[i])  # positions i+2k (odd) -> OD index i+2k-1
    pg.emit("mov", "r0", ev[0])
    for k in range(1, 16):
        pg.emit("add.cc" if k == 1 else ("addc.cc" if k < 15 else "addc"), "r%d" % k, ev[k], od[k - 1])
    return pg


def run_mulwide(pg, x, y):
    env = {}
    for k in range(8):
        env["a%d" % k] = (x >> (32 * k)) & M32
        env["b%d" % k] = (y >> (32 * k)) & M32
    out = pg.run(env)
    return sum(out["r%d" % k] << (32 * k) for k in range(16))


def gen_add(mod):
    pg = Prog()
    t = [pg.new("t") for _ in range(8)]
    for k in range(8):
        pg.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), t[k], "a%d" % k, "b%d" % k)
    cond_sub(pg, t, mod, ["r%d" % k for k in range(8)])
    return pg


def gen_sub(mod):
    pg = Prog()
    pg.allow_wrap = True
    ml = limbs32(mod)
    t = [pg.new("t") for _ in range(8)]
    for k in range(8):
        pg.emit("sub.cc" if k == 0 else "subc.cc", t[k], "a%d" % k, "b%d" % k)
    brw = pg.new("w")
    pg.emit("subc", brw, 0, 0)
    for k in range(8):
        mk = pg.new("k")
        pg.emit("and", mk, brw, ml[k])
        pg.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), "r%d" % k, t[k], mk)
    return pg


def run_binop(pg, x, y):
    env = {}
    for k in range(8):
        env["a%d" % k] = (x >> (32 * k)) & M32
        env["b%d" % k] = (y >> (32 * k)) & M32
    out = pg.run(env)
    return sum(out["r%d" % k] << (32 * k) for k in range(8))


def self_check(trials=300, seed=1):
    rnd = random.Random(seed)
    Rinv = {m: pow(1 << 256, -1, m) for m in (P, R)}
    for mod in (P, R):
        pm, ps, pa, pb = gen_mul(mod), gen_sqr_sos(mod), gen_add(mod), gen_sub(mod)
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253), (1 << 254) % mod, M32, (1 << 224) - 1, mod >> 1]
        vals = edge + [rnd.randrange(mod) for _ in range(trials)]
        for i, x in enumerate(vals):
            y = vals[(i * 7 + 3) % len(vals)]
            assert run_binop(pm, x, y) == x * y * Rinv[mod] % mod, ("mul", hex(x), hex(y))
            assert run_binop(ps, x, 0) == x * x * Rinv[mod] % mod, ("sqr", hex(x))
            assert run_binop(pa, x, y) == (x + y) % mod, ("add", hex(x), hex(y))
            assert run_binop(pb, x, y) == (x - y) % mod, ("sub", hex(x), hex(y))
        pk = gen_mul_karatsuba(mod)
        half = [(1 << 128) - 1, 1 << 128, (1 << 128) + 1, ((1 << 125) - 1) << 128 | ((1 << 128) - 1), ((1 << 128) - 1) << 64]
        kvals = vals + [h % mod for h in half]
        for i, x in enumerate(kvals):
            y = kvals[(i * 7 + 3) % len(kvals)]
            assert run_binop(pk, x, y) == x * y * Rinv[mod] % mod, ("mulk", hex(x), hex(y))
        for x in half:
            for y in half:
                assert run_binop(pk, x % mod, y % mod) == (x % mod) * (y % mod) * Rinv[mod] % mod
        p2 = gen_mul2(mod)
        for i, x in enumerate(vals):
            y, z, w = vals[(i * 7 + 3) % len(vals)], vals[(i * 11 + 5) % len(vals)], vals[(i * 13 + 1) % len(vals)]
            assert run_quadop(p2, x, y, z, w) == (x * y + z * w) * Rinv[mod] % mod, ("mul2", hex(x), hex(y), hex(z), hex(w))
        for x in edge[3:5]:  # both products at their maximum
            assert run_quadop(p2, x, x, x, x) == 2 * x * x * Rinv[mod] % mod
    pw = gen_mulwide()
    full = (1 << 256) - 1
    wvals = [0, 1, full, full - 1, 1 << 255, M32, (1 << 224) - 1, R - 1, P - 1] + [rnd.randrange(1 << 256) for _ in range(trials)]
    for i, x in enumerate(wvals):
        y = wvals[(i * 5 + 2) % len(wvals)]
        assert run_mulwide(pw, x, y) == x * y, ("mulwide", hex(x), hex(y))
    assert run_mulwide(pw, full, full) == full * full
    return True


def emit_fn(name, pg, nin):
    names = {}
    idx = 0
    for k in range(8):
        names["r%d" % k] = "%%%d" % idx
        idx += 1
    for k in range(8):
        names["a%d" % k] = "%%%d" % idx
        idx += 1
    for pre in {1: "", 2: "b", 4: "bcd"}[nin]:
        for k in range(8):
            names["%s%d" % (pre, k)] = "%%%d" % idx
            idx += 1
    lines = pg.ptx(names)
    body = "\n".join('        "%s\\n\\t"' % ln for ln in lines)
    outs = ", ".join('"=r"(r[%d])' % k for k in range(8))
    ins = ", ".join('"r"(a[%d])' % k for k in range(8))
    sig = "const uint32_t (&a)[8]"
    for pre in {1: "", 2: "b", 4: "bcd"}[nin]:
        ins += ", " + ", ".join('"r"(%s[%d])' % (pre, k) for k in range(8))
        sig += ", const uint32_t (&%s)[8]" % pre
    nmul = sum(1 for i in pg.ins if i[0].startswith(("mul", "mad")))
    return ("// %d multiply(-add) PTX instructions; lo/hi pairs fuse into IMAD.WIDE.U32\n"
            "__device__ __forceinline__ void %s(uint32_t (&r)[8], %s) {\n    asm(\n%s\n        : %s\n        : %s);\n}\n"
            % (nmul, name, sig, body, outs, ins))


def emit_fn_wide(name, pg):
    names = {}
    idx = 0
    for k in range(16):
        names["r%d" % k] = "%%%d" % idx
        idx += 1
    for pre in ("a", "b"):
        for k in range(8):
            names["%s%d" % (pre, k)] = "%%%d" % idx
            idx += 1
    lines = pg.ptx(names)
    body = "\n".join('        "%s\\n\\t"' % ln for ln in lines)
    outs = ", ".join('"=r"(r[%d])' % k for k in range(16))
    ins = ", ".join('"r"(a[%d])' % k for k in range(8)) + ", " + ", ".join('"r"(b[%d])' % k for k in range(8))
    nmul = sum(1 for i in pg.ins if i[0] == "mul.wide" or (i[0].startswith("mad") and ".lo" in i[0]))
    return ("// %d wide multiply(-add)s: the 512-bit product a * b, not reduced\n"
            "__device__ __forceinline__ void %s(uint32_t (&r)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {\n"
            "    asm(\n%s\n        : %s\n        : %s);\n}\n" % (nmul, name, body, outs, ins))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "co-zkvms_b200", "csrc", "field_ptx.inc"))
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    self_check(200 if not args.check else 2000)
    parts = ["// GENERATED by tools/gen_field_ptx.py - do not edit.  8 x 32-bit limbs, little-endian, Montgomery R = 2^256.\n"
             "// Every function returns a fully reduced value in [0, modulus).\n"]
    parts.append(emit_fn("fq_mul_ptx", gen_mul(P), 2))
    parts.append(emit_fn("fq_sqr_ptx", gen_sqr_sos(P), 1))
    parts.append(emit_fn("fq_mul2_ptx", gen_mul2(P), 4))
    parts.append(emit_fn("fq_add_ptx", gen_add(P), 2))
    parts.append(emit_fn("fq_sub_ptx", gen_sub(P), 2))
    parts.append(emit_fn("fr_mul_ptx", gen_mul(R), 2))
    parts.append(emit_fn("fr_add_ptx", gen_add(R), 2))
    parts.append(emit_fn("fr_sub_ptx", gen_sub(R), 2))
    parts.append(emit_fn_wide("mulwide_ptx", gen_mulwide()))
    with open(args.o, "w") as f:
        f.write("\n".join(parts))
    print("wrote", os.path.normpath(args.o))


if __name__ == "__main__":
    main()
