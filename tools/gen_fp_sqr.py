#!/usr/bin/env python3
"""Generates curdleproofs_b200/csrc/fp_sqr_rows.inc: the body of the row-wise Montgomery squaring `fp_sqr_rw`.

    T = a^2 (24 words):  off-diagonal products a_i a_j (i < j), row by row, into two arrays -- E takes the products whose word
        position i + j is even, O the odd ones -- so that every row is two plain carry chains of IMAD.WIDE (no third accumulator
        word: the column-wise squaring spends one IADD3.X per product on it); S = E + O; doubled with 23 funnel shifts; the twelve
        diagonal squares added in one 24-word chain.
    reduction: twelve even/odd Montgomery rows on the low half (the same macros as fp_mul_eo, with p * m in both chains), the
        high half added at the end.
    66 + 12 + 144 = 222 wide multiply-adds and ~160 other instructions (column-wise: 234 and ~380).

The script also runs the exact instruction sequence on a 32-bit register / carry-flag model against Python big integers
(random and extreme inputs), so the generated file is checked before it is ever compiled:  python tools/gen_fp_sqr.py
"""
import os
import random

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
INV32 = 0xFFFCFFFD
M32 = 0xFFFFFFFF
PL = [(P >> (32 * i)) & M32 for i in range(12)]
N = 12


# ---------------------------------------------------------------- a tiny PTX-like model
class Machine:
    def __init__(self, a):
        self.r = {f"a{i}": (a >> (32 * i)) & M32 for i in range(N)}
        for i in range(N):
            self.r[f"p{i}"] = PL[i]
        self.cc = 0

    def val(self, x):
        return x if isinstance(x, int) else self.r[x]

    def run(self, ops):
        for op in ops:
            name, d, *s = op
            v = [self.val(x) for x in s]
            cin = self.cc if name.startswith(("madc", "addc")) else 0
            if name.startswith("mad"):
                prod = v[0] * v[1]
                part = (prod & M32) if ".lo" in name else (prod >> 32)
                t = part + v[2] + cin
            elif name.startswith("add"):
                t = v[0] + v[1] + cin
            elif name == "shf":      # d = (hi << 1) | (lo >> 31)
                t = ((v[1] << 1) | (v[0] >> 31)) & M32
            elif name == "mullo":
                t = (v[0] * v[1]) & M32
            elif name == "mov":
                t = v[0]
            elif name == "square_done":
                self.square = sum(self.r[f"E{k}"] << (32 * k) for k in range(24))
                continue
            else:
                raise ValueError(name)
            if name.endswith(".cc"):
                self.cc = t >> 32
            else:
                assert t >> 32 == 0 or name in ("mullo",), (op, hex(t))  # a dropped carry would be a bug
            self.r[d] = t & M32


# ---------------------------------------------------------------- the schedule
def offdiag_chains():
    """[(array, first word, [j...], multiplier i)] for rows i = 1..10; row 0 is plain products."""
    out = []
    for i in range(1, N - 1):
        odd = [j for j in range(i + 1, N) if (i + j) % 2 == 1]
        even = [j for j in range(i + 1, N) if (i + j) % 2 == 0]
        if odd:
            out.append(("O", i + odd[0], odd, i))
        if even:
            out.append(("E", i + even[0], even, i))
    return out


def chain_ops(arr, w0, js, i, hi):
    """Instruction list of one accumulate chain; `hi[arr]` = highest initialised word of the array (updated)."""
    ops = []
    L = len(js)
    last_fresh = False
    for k, j in enumerate(js):
        for half, part in enumerate(("lo", "hi")):
            w = w0 + 2 * k + half
            fresh = w > hi[arr]
            first = k == 0 and half == 0
            name = ("mad." if first else "madc.") + part + ".cc"
            ops.append((name, f"{arr}{w}", f"a{j}", f"a{i}", 0 if fresh else f"{arr}{w}"))
            last_fresh = fresh
    top = w0 + 2 * L
    if last_fresh:
        # the top word of the chain was fresh: product high word + carry cannot overflow, the chain simply ends
        n, d, *s = ops[-1]
        ops[-1] = (n[:-3], d, *s)
        hi[arr] = max(hi[arr], top - 1)
    else:
        assert top > hi[arr], "carry would land in a live word"
        ops.append(("addc", f"{arr}{top}", 0, 0))
        hi[arr] = top
    return ops


def mad6_ops(acc, mul, s, top):
    """CDP_MAD6(acc, m0..m5, s, top): acc[0..11] += (m_k * s) at words 2k, 2k+1; top += carry."""
    ops = []
    for k in range(6):
        ops.append((("mad." if k == 0 else "madc.") + "lo.cc", acc[2 * k], mul[k], s, acc[2 * k]))
        ops.append(("madc.hi.cc", acc[2 * k + 1], mul[k], s, acc[2 * k + 1]))
    ops.append(("addc", top, top, 0))
    return ops


def mad6_rshift_ops(X, y0, mul, s):
    """CDP_MAD6_RSHIFT(X, y0, m0..m5, s): y0 += X[1]; X <- (X >> 64) + m * s, in place."""
    ops = [("add.cc", y0, y0, X[1])]
    for k in range(6):
        lo_add = X[2 * k + 2] if 2 * k + 2 < 12 else 0
        hi_add = X[2 * k + 3] if 2 * k + 3 < 12 else 0
        ops.append(("madc.lo.cc", X[2 * k], mul[k], s, lo_add))
        ops.append((("madc.hi.cc" if k < 5 else "madc.hi"), X[2 * k + 1], mul[k], s, hi_add))
    return ops


def schedule():
    """The whole squaring as (simulator ops, C text)."""
    ops, c = [], []
    hi = {"E": 11, "O": 12}
    # row 0: plain products
    for j in range(1, N):
        arr = "O" if j % 2 else "E"
        ops += [("mad.lo", f"{arr}{j}", f"a{j}", "a0", 0), ("mad.hi", f"{arr}{j + 1}", f"a{j}", "a0", 0)]
    c.append("    // row 0: plain products a_j * a_0 (odd j -> O, even j -> E)")
    c.append("#pragma unroll\n    for (int j = 1; j < 12; j++) {\n        const uint64_t pr = (uint64_t)a.v[j] * a.v[0];\n"
             "        if (j & 1) { O[j] = (uint32_t)pr; O[j + 1] = (uint32_t)(pr >> 32); } else { E[j] = (uint32_t)pr; E[j + 1] = (uint32_t)(pr >> 32); }\n    }")
    # rows 1..10
    for arr, w0, js, i in offdiag_chains():
        co = chain_ops(arr, w0, js, i, hi)
        ops += co
        c.append(emit_asm(co, f"row {i}, {'odd' if arr == 'O' else 'even'} positions: a[{js[0]}..] * a[{i}] into {arr}[{w0}..]"))
    assert hi == {"E": 21, "O": 22}, hi
    # S = E + O (E lives in words 2..21, O in 1..22), kept in O; then doubled into T (E registers), T[0] = 0
    merge = [("add.cc", "O2", "O2", "E2")] + [("addc.cc", f"O{k}", f"O{k}", f"E{k}") for k in range(3, 22)] + [("addc", "O22", "O22", 0)]
    ops += merge
    c.append(emit_asm(merge, "S = E + O (E: words 2..21, O: words 1..22)"))
    dbl = [("shf", f"E{k}", f"O{k - 1}" if k >= 2 else 0, f"O{k}" if k <= 22 else 0) for k in range(23, 0, -1)]
    ops += dbl
    c.append("    // T = 2 S: one funnel shift per word (no carry chain)")
    c.append("#pragma unroll\n    for (int k = 23; k >= 1; k--) E[k] = __funnelshift_l(k >= 2 ? O[k - 1] : 0u, k <= 22 ? O[k] : 0u, 1);\n    E[0] = 0;")
    ops.append(("mov", "E0", 0))
    # diagonal: T += sum a_i^2 2^(64 i), one chain over the 24 words
    diag = []
    for i in range(N):
        diag.append((("mad." if i == 0 else "madc.") + "lo.cc", f"E{2 * i}", f"a{i}", f"a{i}", f"E{2 * i}"))
        diag.append(("madc.hi.cc" if i < N - 1 else "madc.hi", f"E{2 * i + 1}", f"a{i}", f"a{i}", f"E{2 * i + 1}"))
    ops += diag
    c.append(emit_asm(diag, "diagonal squares"))
    ops.append(("square_done", "", 0))
    # reduction of the low half: X = T[0..11] aligned, Y = 0 offset by one word
    X = [f"E{k}" for k in range(12)]
    Y = [f"Y{k}" for k in range(12)]
    ops += [("mov", y, 0) for y in Y]
    podd, peven = [f"p{k}" for k in range(1, 12, 2)], [f"p{k}" for k in range(0, 12, 2)]
    # row 0 = CDP_REDC_ROW(X, Y)
    ops.append(("mullo", "m", X[0], INV32))
    ops.append(("mov", "drop", 0))
    ops += mad6_ops(Y, podd, "m", "drop")
    ops += mad6_ops(X, peven, "m", Y[11])
    A, B = X, Y   # A: aligned (word 0 now zero), B: offset
    for _ in range(1, 12):
        # m = new aligned word 0 = B[0] + A[1]
        ops.append(("add.cc", "t", B[0], A[1]))   # value only; the carry is recomputed inside the shift chain
        ops.append(("mullo", "m", "t", INV32))
        ops += mad6_rshift_ops(A, B[0], podd, "m")
        ops += mad6_ops(B, peven, "m", A[11])
        A, B = B, A
    # now A = aligned (word 0 zero), B = offset: U = B + (A >> 32); result = U + T[12..23]
    fin = [("add.cc", "r0", B[0], A[1])] + [("addc.cc", f"r{k}", B[k], A[k + 1]) for k in range(1, 11)] + [("addc", "r11", B[11], 0)]
    fin += [("add.cc", "r0", "r0", "E12")] + [("addc.cc", f"r{k}", f"r{k}", f"E{12 + k}") for k in range(1, 11)] + [("addc", "r11", "r11", "E23")]
    ops += fin
    final_aligned_is_X = A is X
    return ops, c, final_aligned_is_X


def emit_asm(ops, comment):
    """One asm statement for a carry chain.  Registers named E<k>/O<k> map to E[k]/O[k], a<k> to a.v[k]."""
    outs, ins = [], []

    def cexpr(x):
        return f"{x[0]}[{x[1:]}]" if x[0] in "EO" else f"a.v[{x[1:]}]"

    written = []
    for name, d, *s in ops:
        if d not in written:
            written.append(d)
    # a destination that is also read as an addend is "+r", otherwise "=r"
    read_before_write = set()
    seen = set()
    for name, d, *s in ops:
        for x in s:
            if isinstance(x, str) and x[0] in "EO" and x not in seen:
                read_before_write.add(x)
        seen.add(d)
    for d in written:
        outs.append((d, "+r" if d in read_before_write else "=r"))
    for name, d, *s in ops:
        for x in s:
            if isinstance(x, str) and x not in [o[0] for o in outs] and x not in ins:
                ins.append(x)
    idx = {o[0]: k for k, o in enumerate(outs)}
    for x in ins:
        idx[x] = len(idx)
    lines = []
    for name, d, *s in ops:
        def o(x):
            return "0" if x == 0 else f"%{idx[x]}"
        if name.startswith("mad"):
            lines.append(f"{name}.u32 {o(d)}, {o(s[0])}, {o(s[1])}, {o(s[2])};")
        else:
            lines.append(f"{name}.u32 {o(d)}, {o(s[0])}, {o(s[1])};")
    body = "\\n\\t\"\n        \"".join(lines)
    outs_s = ", ".join(f'"{c}"({cexpr(d)})' for d, c in outs)
    ins_s = ", ".join(f'"r"({cexpr(x)})' for x in ins)
    return f"    // {comment}\n    asm(\"{body}\"\n        : {outs_s}\n        : {ins_s});"


def simulate(a):
    ops, _, _ = schedule()
    m = Machine(a)
    m.run(ops)
    assert m.r["drop"] == 0 and m.square == a * a
    return sum(m.r[f"r{k}"] << (32 * k) for k in range(12))


def main():
    rnd = random.Random(1)
    Rinv = pow(1 << 384, -1, P)
    cases = [0, 1, P - 1, P - 2, (1 << 381) - 1, int("ffffffff" * 12, 16) % P, int("ffffffff00000000" * 6, 16) % P, int("00000000ffffffff" * 6, 16)] + \
            [rnd.randrange(P) for _ in range(3000)]
    # the unreduced sum is only promised for inputs < p, but the carry model must also hold for the largest 384-bit patterns
    for a in cases:
        got = simulate(a)
        want = a * a * Rinv % P
        assert got % P == want and got < 2 * P, (hex(a), hex(got), hex(want))
    for a in (int("ffffffff" * 12, 16), int("ffffffff" * 11 + "fffffffe", 16), int("ffffffff00000000" * 6, 16), 1 << 383):
        # not field elements: the 24-word square itself must still be exact for every 384-bit pattern (no dropped carry)
        ops, _, _ = schedule()
        m = Machine(a)
        m.run(ops[:next(k for k, o in enumerate(ops) if o[0] == "square_done") + 1])
        assert m.square == a * a, hex(a)
    ops, c, final_aligned_is_X = schedule()
    wide = sum(1 for o in ops if o[0].startswith("mad") and ".lo" in o[0])
    print(f"model ok on {len(cases)} inputs; {wide} wide multiply-adds, {len(ops) - 2 * wide - 14} other instructions (model count)")
    assert not final_aligned_is_X
    out = ["// GENERATED by tools/gen_fp_sqr.py -- do not edit.  Body of fp_sqr_rw(fp &r, const fp &a): see the generator for the schedule.",
           "    uint32_t E[24], O[24];"]
    out += c
    out.append("    // Montgomery reduction of T[0..11] (even/odd rows, as in fp_mul_eo), then the high half")
    out.append("    uint32_t Y[12];\n#pragma unroll\n    for (int k = 0; k < 12; k++) Y[k] = 0;")
    out.append("    CDP_REDC_ROW(E, Y)")
    out.append("    CDP_SQR_RED_ROW(E, Y) CDP_SQR_RED_ROW(Y, E) CDP_SQR_RED_ROW(E, Y) CDP_SQR_RED_ROW(Y, E) CDP_SQR_RED_ROW(E, Y) CDP_SQR_RED_ROW(Y, E)")
    out.append("    CDP_SQR_RED_ROW(E, Y) CDP_SQR_RED_ROW(Y, E) CDP_SQR_RED_ROW(E, Y) CDP_SQR_RED_ROW(Y, E) CDP_SQR_RED_ROW(E, Y)")
    out.append("    // aligned array: Y (word 0 zero), offset array: E[0..11]; result = E + (Y >> 32) + T[12..23]")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curdleproofs_b200", "csrc", "fp_sqr_rows.inc")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
