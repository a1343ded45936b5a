#!/usr/bin/env python3
"""Generate `fq_unsat.cuh`: carry-free Fq arithmetic for the MSM hot loop.

Why: on B200 the plain `IMAD.WIDE.U32` issues at 54 lanes/clk/SM, but every form that reads or
writes the carry predicate (`IMAD.WIDE.U32.X`, what a saturated 32-bit-limb Montgomery product is
made of) issues at ~31 (profiles/r01_imad_variants.txt).  So the hot loop works on UNSATURATED limbs:

  * a value is 13 limbs of 30 bits (`fqu_t`, 390 bits of room for a 381-bit modulus);
  * products are accumulated in 64-bit registers with plain `mad.wide.u32` -- 13 terms of < 2^60
    never wrap -- and carries are moved with shifts/adds on the ALU pipe, which is otherwise idle;
  * Montgomery radix is R' = 2^390.  Wire values (R = 2^384) enter by a 6-bit shift (x64, no
    reduction needed: 64p < 2^390) and leave through one product with 2^384;
  * reduction is lazy: products return a value < 1.1p, linear combinations carry a small multiple of
    p as bias and are renormalised; nothing is compared with p inside the loop.

Multiplication = separated operand scanning: 13 rows of the product (columns normalised as they
complete), then 13 reduction rows, 169 + 169 (+13 for the m_i) wide multiply-adds.

Every routine is a `ptxir.Block`: printed as ONE inline-PTX block and executed in Python against big
integers with overflow checking on every 64-bit accumulate (tests/test_gen_unsat.py).
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "midnight_bls12_381_cuda_b200", "csrc", "gen"))
from ptxir import Block, M32  # noqa: E402

BLS_X = -0xD201000000010000
R_MOD = BLS_X**4 - BLS_X**2 + 1
P = ((BLS_X - 1) ** 2 * R_MOD) // 3 + BLS_X

W = 30
L = 13
MASK = (1 << W) - 1
RP = 1 << (W * L)                        # R' = 2^390
M0 = (-pow(P, -1, 1 << W)) % (1 << W)    # -p^-1 mod 2^30
WIRE_SHIFT = W * L - 384                 # 6


def limbs(v: int) -> list[int]:
    return [(v >> (W * i)) & MASK for i in range(L)]


def value(ls) -> int:
    return sum(int(x) << (W * i) for i, x in enumerate(ls))


P_L = limbs(P)
ONE_INT = RP % P                          # 1 in the internal Montgomery domain
TO_WIRE = (1 << 384) % P                  # x_int (*) TO_WIRE = x * 2^384  (wire Montgomery form)


def normalize64(b: Block, acc: list, out: list, n: int):
    """acc[0..n-1] 64-bit column sums -> 30-bit limbs (appended to out); returns final carry (64-bit)."""
    c = None
    for s in range(n):
        a = acc[s] if c is None else b.w3("add.u64", acc[s], c)
        lo = b.op2("cvt.u32.u64", a)
        out.append(b.op3("and.b32", lo, MASK))
        c = b.w3("shr.u64", a, W)
    return c


def product_rows(b: Block, a: list, bb: list) -> list:
    """26 normalised limbs of a*b: 13 rows over a sliding window of 13 64-bit column sums."""
    T: list = []
    acc: list = [None] * L
    for i in range(L):
        for s in range(L):
            acc[s] = b.w3("mul.wide.u32", a[s], bb[i]) if acc[s] is None else b.w4("mad.wide.u32", a[s], bb[i], acc[s])
        # column i is complete: emit its limb, push the carry into the next column, slide the window
        lo = b.op2("cvt.u32.u64", acc[0])
        T.append(b.op3("and.b32", lo, MASK))
        c = b.w3("shr.u64", acc[0], W)
        acc = acc[1:] + [None]
        acc[0] = b.w3("add.u64", acc[0], c)
    # acc[0..L-2] hold columns L .. 2L-2
    c = normalize64(b, acc[: L - 1], T, L - 1)
    T.append(b.op2("cvt.u32.u64", c))          # column 2L-1
    return T


def product_rows_square(b: Block, a: list) -> list:
    """a^2 with cross terms taken once against 2a: 91 wide products instead of 169."""
    T: list = []
    a2 = [b.op3("shl.b32", x, 1) for x in a]
    col: list = [None] * (2 * L)               # absolute columns

    def add(k, x, y):
        col[k] = b.w3("mul.wide.u32", x, y) if col[k] is None else b.w4("mad.wide.u32", x, y, col[k])

    carry = None
    for i in range(L):
        add(2 * i, a[i], a[i])
        for j in range(i + 1, L):
            add(i + j, a[i], a2[j])
        # column i complete (all its rows i' <= i/2 done)
        head = col[i]
        if carry is not None:
            head = b.w3("add.u64", head, carry)
        lo = b.op2("cvt.u32.u64", head)
        T.append(b.op3("and.b32", lo, MASK))
        carry = b.w3("shr.u64", head, W)
    for k in range(L, 2 * L - 1):
        head = b.w3("add.u64", col[k], carry)
        lo = b.op2("cvt.u32.u64", head)
        T.append(b.op3("and.b32", lo, MASK))
        carry = b.w3("shr.u64", head, W)
    T.append(b.op2("cvt.u32.u64", carry))
    return T


def montgomery_reduce(b: Block, T: list, outs: list[str]):
    """(T + M p) / 2^390 from 26 normalised limbs; result normalised, < T/2^390 + p."""
    acc = [b.w2("cvt.u64.u32", T[s]) for s in range(L)]
    for i in range(L):
        lo = b.op2("cvt.u32.u64", acc[0])
        m = b.op3("and.b32", b.op3("mul.lo.u32", lo, M0), MASK)
        for s in range(L):
            acc[s] = b.w4("mad.wide.u32", m, P_L[s], acc[s])
        c = b.w3("shr.u64", acc[0], W)
        nxt = b.w2("cvt.u64.u32", T[L + i])
        acc = acc[1:] + [nxt]
        acc[0] = b.w3("add.u64", acc[0], c)
    c = None
    for s in range(L):
        a = acc[s] if c is None else b.w3("add.u64", acc[s], c)
        lo = b.op2("cvt.u32.u64", a)
        b.op3("and.b32", lo, MASK, dst=outs[s])
        c = b.w3("shr.u64", a, W)


def mul_body(b: Block, a, bb, outs):
    montgomery_reduce(b, product_rows(b, a, bb), outs)


def sqr_body(b: Block, a, outs):
    montgomery_reduce(b, product_rows_square(b, a), outs)


def lincomb_body(b: Block, terms, K: int, outs):
    """outs = sum coef*x + K*p, renormalised.  coefs in {+1,-1,+2,-2}; limbs treated as signed 32-bit
    during the carry sweep (arithmetic shift), the bias keeps the total positive."""
    KP = limbs(K * P)
    s = []
    for k in range(L):
        cur = KP[k]
        for coef, x in terms:
            for _ in range(abs(coef)):
                cur = b.op3("add.u32" if coef > 0 else "sub.u32", cur, x[k])
        s.append(cur)
    for k in range(L - 1):
        c = b.op3("shr.s32", s[k], W)
        b.op3("and.b32", s[k], MASK, dst=outs[k])
        s[k + 1] = b.op3("add.u32", s[k + 1], c)
    b.mov(s[L - 1], dst=outs[L - 1])


def from_wire_body(b: Block, w, outs):
    """12 saturated 32-bit words (wire Montgomery form, value V < 2^384) -> limbs of V * 2^6."""
    ww = list(w) + [0]
    for k in range(L):
        if k == 0:
            t = b.op3("shl.b32", ww[0], WIRE_SHIFT)
            b.op3("and.b32", t, MASK, dst=outs[0])
            continue
        start = W * k - WIRE_SHIFT
        q, o = start // 32, start % 32
        hi = ww[q + 1] if q + 1 < len(ww) else 0
        x = b.op4("shf.r.wrap.b32", ww[q], hi, o) if o else ww[q]
        b.op3("and.b32", x, MASK, dst=outs[k])


def to_words_canonical_body(b: Block, u, outs12):
    """normalised value < 2p (already in wire scale) -> canonical (< p) packed into 12 32-bit words."""
    # d = u - p with a signed carry sweep; the sign of the top limb says whether u >= p
    d = [b.op3("sub.u32", u[k], P_L[k]) for k in range(L)]
    dn = []
    for k in range(L - 1):
        c = b.op3("shr.s32", d[k], W)
        dn.append(b.op3("and.b32", d[k], MASK))
        d[k + 1] = b.op3("add.u32", d[k + 1], c)
    dn.append(d[L - 1])
    neg = b.op3("shr.s32", d[L - 1], 31)            # all-ones if u < p
    sel = [b.mask_select(neg, u[k], dn[k]) for k in range(L)]
    # pack 13 x 30 bits -> 12 x 32 bits
    for j in range(12):
        bit = 32 * j
        q, o = bit // W, bit % W
        lo = b.op3("shr.u32", sel[q], o) if o else sel[q]
        have = W - o
        word = lo
        if have < 32 and q + 1 < L:
            hi = b.op3("shl.b32", sel[q + 1], have)
            word = b.op3("or.b32", lo, hi)
            have += W
            if have < 32 and q + 2 < L:
                hi2 = b.op3("shl.b32", sel[q + 2], have)
                word = b.op3("or.b32", word, hi2)
        b.mov(word, dst=outs12[j])


def zero_test_body(b: Block, u, outs2):
    """outs2[0] == 0 <=> all limbs zero; outs2[1] == 0 <=> limbs == p  (u normalised, < 2p)."""
    z = u[0]
    for k in range(1, L):
        z = b.op3("or.b32", z, u[k])
    e = b.op3("xor.b32", u[0], P_L[0])
    for k in range(1, L):
        e = b.op3("or.b32", e, b.op3("xor.b32", u[k], P_L[k]))
    b.mov(z, dst=outs2[0])
    b.mov(e, dst=outs2[1])


# ---------------------------------------------------------------------------
ROUTINES = {
    # name: (inputs [(prefix, count)], n_out, builder)
    "fqu_mul": ([("a", L), ("b", L)], L, lambda b, i, o: mul_body(b, i[0], i[1], o)),
    "fqu_sqr": ([("a", L)], L, lambda b, i, o: sqr_body(b, i[0], o)),
    "fqu_sub_k2": ([("a", L), ("b", L)], L, lambda b, i, o: lincomb_body(b, [(1, i[0]), (-1, i[1])], 2, o)),
    "fqu_sub_k4": ([("a", L), ("b", L)], L, lambda b, i, o: lincomb_body(b, [(1, i[0]), (-1, i[1])], 4, o)),
    "fqu_sub_k8": ([("a", L), ("b", L)], L, lambda b, i, o: lincomb_body(b, [(1, i[0]), (-1, i[1])], 8, o)),
    # a - 2b + 4p.  (A three-term a - b - 2c does not fit signed 32-bit limbs: its span is 2^32.)
    "fqu_sub2_k4": ([("a", L), ("b", L)], L, lambda b, i, o: lincomb_body(b, [(1, i[0]), (-2, i[1])], 4, o)),
    "fqu_neg_k2": ([("a", L)], L, lambda b, i, o: lincomb_body(b, [(-1, i[0])], 2, o)),
    "fqu_from_wire": ([("a", 12)], L, lambda b, i, o: from_wire_body(b, i[0], o)),
    "fqu_pack_canonical": ([("a", L)], 12, lambda b, i, o: to_words_canonical_body(b, i[0], o)),
    "fqu_zero_test": ([("a", L)], 2, lambda b, i, o: zero_test_body(b, i[0], o)),
}


def build(name: str) -> Block:
    ins, nout, fn = ROUTINES[name]
    b = Block(name)
    regs = [[b.inp(f"{p}{k}") for k in range(n)] for p, n in ins]
    outs = [f"r{k}" for k in range(nout)]
    b.outputs = outs
    fn(b, regs, outs)
    return b


def run(name: str, *vals) -> list[int]:
    """vals: one list of limb/word ints per input; returns output regs."""
    ins, nout, _ = ROUTINES[name]
    blk = build(name)
    env = {}
    for (p, n), v in zip(ins, vals):
        assert len(v) == n
        for k, x in enumerate(v):
            env[f"{p}{k}"] = x & M32
    r = blk.run(env, strict64=True)
    return [r[f"r{k}"] for k in range(nout)]


def emit_cpp(name: str) -> str:
    ins, nout, _ = ROUTINES[name]
    blk = build(name)
    lines = blk.ptx_lines()
    decl = [f".reg .u32 t<{max(blk.nreg, 1)}>;"]
    if getattr(blk, "nreg64", 0):
        decl.append(f".reg .u64 d<{blk.nreg64}>;")
    decl.append(".reg .u32 " + ", ".join(f"{p}<{n}>" for p, n in ins) + f", r<{nout}>;")
    if getattr(blk, "_npredsel", 0):
        decl.append(f".reg .pred q<{blk._npredsel}>;")
    body = ["{"] + decl
    idx = nout
    for p, n in ins:
        for k in range(n):
            body.append(f"mov.u32 {p}{k}, %{idx};")
            idx += 1
    body += lines
    for k in range(nout):
        body.append(f"mov.u32 %{k}, r{k};")
    body.append("}")
    text = "\n".join(f'      "{l}\\n\\t"' for l in body)
    params, operands = [], []
    for p, n in ins:
        params.append(f"const uint32_t (&{p})[{n}]")
        operands += [f'"r"({p}[{k}])' for k in range(n)]
    outs = ", ".join(f'"=r"(r[{k}])' for k in range(nout))
    # outputs are written only after every input has been copied into block-local registers
    return (f"__device__ __forceinline__ void {name}_raw(uint32_t (&r)[{nout}], {', '.join(params)}) {{\n"
            f"  asm(\n{text}\n      : {outs}\n      : {', '.join(operands)});\n}}\n")


HEADER = """// GENERATED by csrc/gen/gen_unsat.py -- do not edit.
// Carry-free (unsaturated 13 x 30-bit limb) Fq arithmetic for the MSM hot loop; Montgomery radix 2^390.
// See the generator's docstring for the design and the bounds.
#pragma once
#include <cstdint>

"""


def constants() -> str:
    def arr(v):
        return "{" + ", ".join(f"0x{x:08x}u" for x in limbs(v)) + "}"
    return (f"#define FQU_LIMBS {L}\n#define FQU_ONE_INIT {arr(ONE_INT)}      /* 2^390 mod p */\n"
            f"#define FQU_TO_WIRE_INIT {arr(TO_WIRE)}  /* 2^384 mod p: x_int (*) this = x in wire Montgomery form */\n\n")


def generate() -> str:
    return HEADER + constants() + "\n".join(emit_cpp(n) for n in ROUTINES)


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fq_unsat.cuh")
    with open(dst, "w") as fh:
        fh.write(generate())
    for n in ROUTINES:
        print(n, build(n).count())
