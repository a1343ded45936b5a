"""Tiny straight-line PTX IR with two back ends: PTX text (for nvcc inline asm)
and a Python interpreter (so every generated routine is checked against big
integers on the CPU before it ever reaches a GPU).

Only the u32 carry-chain subset that the field arithmetic needs is modelled:
    mul.lo/hi, mad.lo/hi[.cc], madc.lo/hi[.cc], add[.cc], addc[.cc],
    sub[.cc], subc[.cc], mov, and, selp-by-mask (setp.ne + selp), shf.
The condition-code semantics follow the PTX ISA: *.cc writes CC.CF,
addc/subc/madc read it; for sub, CF is the *borrow*.
"""
from __future__ import annotations

M32 = 0xFFFFFFFF


class Reg(str):
    pass


class Block:
    def __init__(self, name: str):
        self.name = name
        self.ins: list[tuple] = []
        self.nreg = 0
        self.npred = 0
        self.inputs: list[str] = []    # names of 32-bit input regs
        self.outputs: list[str] = []   # names of 32-bit output regs

    # ---- registers ----
    def reg(self) -> Reg:
        self.nreg += 1
        return Reg(f"t{self.nreg - 1}")

    def reg64(self) -> Reg:
        self.nreg64 = getattr(self, "nreg64", 0) + 1
        return Reg(f"d{self.nreg64 - 1}")

    def pred(self) -> Reg:
        self.npred += 1
        return Reg(f"p{self.npred - 1}")

    def inp(self, name: str) -> Reg:
        self.inputs.append(name)
        return Reg(name)

    # ---- instructions; every emitter returns its destination ----
    def _e(self, op, dst, *src):
        self.ins.append((op, dst, src))
        return dst

    def op3(self, op, a, b, dst=None):
        return self._e(op, dst or self.reg(), a, b)

    def op4(self, op, a, b, c, dst=None):
        return self._e(op, dst or self.reg(), a, b, c)

    def op2(self, op, a, dst=None):
        return self._e(op, dst or self.reg(), a)

    # 64-bit destinations
    def w3(self, op, a, b):
        return self._e(op, self.reg64(), a, b)

    def w4(self, op, a, b, c):
        return self._e(op, self.reg64(), a, b, c)

    def w2(self, op, a):
        return self._e(op, self.reg64(), a)

    def mov(self, a, dst=None):
        return self._e("mov.u32", dst or self.reg(), a)

    def mask_select(self, mask, a, b, dst=None):
        """dst = mask != 0 ? a : b  (mask is 0 or 0xffffffff)."""
        return self._e("selm", dst or self.reg(), mask, a, b)

    # ---- PTX text ----
    @staticmethod
    def _s(x):
        if isinstance(x, int):
            return f"0x{x & M32:08x}"
        return x

    def ptx_lines(self) -> list[str]:
        out = []
        pred_of: dict[str, str] = {}
        for op, dst, src in self.ins:
            if op == "selm":
                mask, a, b = src
                if mask not in pred_of:
                    p = f"q{len(pred_of)}"
                    pred_of[mask] = p
                    out.append(f"setp.ne.u32 {p}, {mask}, 0;")
                out.append(f"selp.u32 {dst}, {self._s(a)}, {self._s(b)}, {pred_of[mask]};")
            else:
                out.append(f"{op} {dst}, " + ", ".join(self._s(s) for s in src) + ";")
        self._npredsel = len(pred_of)
        return out

    # ---- interpreter ----
    def run(self, env: dict[str, int], strict64: bool = False) -> dict[str, int]:
        """strict64: raise if a 64-bit accumulate wraps (used to prove the no-overflow bounds)."""
        r = dict(env)
        cf = 0
        M64 = (1 << 64) - 1

        def v(x):
            return (x & M32) if isinstance(x, int) else r[x]

        for op, dst, src in self.ins:
            wide_src = op in ("add.u64", "shr.u64", "cvt.u32.u64")
            if op == "selm" or wide_src:
                s = None
            elif op == "mad.wide.u32":
                s = [v(src[0]), v(src[1])]
            else:
                s = [v(x) for x in src]
            if op == "mov.u32":
                r[dst] = s[0]
            elif op == "mul.lo.u32":
                r[dst] = (s[0] * s[1]) & M32
            elif op == "mul.hi.u32":
                r[dst] = (s[0] * s[1]) >> 32
            elif op in ("mad.lo.u32", "mad.lo.cc.u32", "madc.lo.u32", "madc.lo.cc.u32",
                        "mad.hi.u32", "mad.hi.cc.u32", "madc.hi.u32", "madc.hi.cc.u32"):
                prod = s[0] * s[1]
                part = (prod & M32) if ".lo" in op else (prod >> 32)
                t = part + s[2] + (cf if op.startswith("madc") else 0)
                r[dst] = t & M32
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("add.u32", "add.cc.u32", "addc.u32", "addc.cc.u32"):
                t = s[0] + s[1] + (cf if op.startswith("addc") else 0)
                r[dst] = t & M32
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("sub.u32", "sub.cc.u32", "subc.u32", "subc.cc.u32"):
                t = s[0] - s[1] - (cf if op.startswith("subc") else 0)
                r[dst] = t & M32
                if ".cc" in op:
                    cf = 1 if t < 0 else 0
            elif op == "and.b32":
                r[dst] = s[0] & s[1]
            elif op == "or.b32":
                r[dst] = s[0] | s[1]
            elif op == "xor.b32":
                r[dst] = s[0] ^ s[1]
            elif op == "shf.l.wrap.b32":       # funnel shift left: (hi:lo) << n, upper word
                lo, hi, n = s
                n &= 31
                r[dst] = ((((hi << 32) | lo) << n) >> 32) & M32
            elif op == "shf.r.wrap.b32":       # funnel shift right: lower word of (hi:lo) >> n
                lo, hi, n = s
                n &= 31
                r[dst] = (((hi << 32) | lo) >> n) & M32
            elif op == "shl.b32":
                r[dst] = (s[0] << s[1]) & M32 if s[1] < 32 else 0
            elif op == "shr.u32":
                r[dst] = (s[0] >> s[1]) if s[1] < 32 else 0
            elif op == "selm":
                mask, a, b = src
                r[dst] = v(a) if v(mask) != 0 else v(b)
            elif op == "mul.wide.u32":
                r[dst] = s[0] * s[1]
            elif op == "mad.wide.u32":
                t = s[0] * s[1] + r[src[2]]
                if strict64 and t > M64:
                    raise OverflowError(f"mad.wide wrapped: {dst}")
                r[dst] = t & M64
            elif op == "add.u64":
                a64 = r[src[0]] if not isinstance(src[0], int) else src[0]
                b64 = r[src[1]] if not isinstance(src[1], int) else src[1]
                t = a64 + b64
                if strict64 and t > M64:
                    raise OverflowError(f"add.u64 wrapped: {dst}")
                r[dst] = t & M64
            elif op == "shr.u64":
                r[dst] = r[src[0]] >> (src[1] if isinstance(src[1], int) else r[src[1]])
            elif op == "cvt.u32.u64":
                r[dst] = r[src[0]] & M32
            elif op == "cvt.u64.u32":
                r[dst] = s[0]
            elif op == "shr.s32":
                x = s[0] - (1 << 32) if s[0] >> 31 else s[0]
                r[dst] = (x >> s[1]) & M32
            else:
                raise NotImplementedError(op)
        return r

    def count(self) -> dict[str, int]:
        c: dict[str, int] = {}
        for op, _, _ in self.ins:
            c[op] = c.get(op, 0) + 1
        return c
