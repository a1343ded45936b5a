"""The carry-free (unsaturated-limb) Fq routines of csrc/gen/gen_unsat.py, executed by the PTX
interpreter with 64-bit overflow checking, against big integers -- plus the whole lazy XYZZ mixed
addition of csrc/fq_lazy.cuh replayed routine by routine and compared with affine point addition."""
import os
import random
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
sys.path[:0] = [os.path.dirname(os.path.abspath(__file__)), _ROOT]
import gen_unsat as U  # noqa: E402

from oracle import pyref as P  # noqa: E402

PM = U.P
RPINV = pow(U.RP, -1, PM)


def test_mul_sqr_bounds_and_values():
    rnd = random.Random(5)
    for t in range(40):
        a = rnd.randrange(10 * PM)
        b = rnd.randrange((64 if t % 3 == 0 else 10) * PM)
        r = U.value(U.run("fqu_mul", U.limbs(a), U.limbs(b)))
        assert r % PM == a * b * RPINV % PM and r < 2 * PM
        s = U.value(U.run("fqu_sqr", U.limbs(a)))
        assert s % PM == a * a * RPINV % PM and s < 2 * PM
    # no 64-bit accumulator can wrap for ANY normalised input: all limbs at 2^30 - 1 (strict64 raises otherwise)
    worst = [U.MASK] * U.L
    U.run("fqu_mul", worst, worst)
    U.run("fqu_sqr", worst)


def test_linear_combinations():
    rnd = random.Random(6)
    for _ in range(60):
        a = rnd.randrange(2 * PM)
        for name, k in (("fqu_sub_k2", 2), ("fqu_sub_k4", 4), ("fqu_sub_k8", 8)):
            b = rnd.randrange(k * PM)
            r = U.run(name, U.limbs(a), U.limbs(b))
            assert all(x <= U.MASK for x in r) and U.value(r) == a - b + k * PM
        a4, b2 = rnd.randrange(4 * PM), rnd.randrange(2 * PM)
        r = U.run("fqu_sub2_k4", U.limbs(a4), U.limbs(b2))
        assert all(x <= U.MASK for x in r) and U.value(r) == a4 - 2 * b2 + 4 * PM
        assert U.value(U.run("fqu_neg_k2", U.limbs(a))) == 2 * PM - a
    assert U.value(U.run("fqu_sub_k8", U.limbs(0), U.limbs(8 * PM - 1))) == 1


def test_wire_conversions_and_zero_test():
    rnd = random.Random(7)
    for t in range(40):
        v = rnd.randrange(PM) if t else PM - 1
        words = [(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
        r = U.run("fqu_from_wire", words)
        assert all(x <= U.MASK for x in r) and U.value(r) == v * 64
    for t in range(60):
        v = rnd.randrange(2 * PM) if t > 3 else [0, PM - 1, PM, 2 * PM - 1][t]
        w = U.run("fqu_pack_canonical", U.limbs(v))
        assert sum(x << (32 * i) for i, x in enumerate(w)) == v % PM
    for v, exp in ((0, True), (PM, True), (1, False), (PM - 1, False), (PM + 1, False)):
        z, e = U.run("fqu_zero_test", U.limbs(v))
        assert ((z == 0) or (e == 0)) == exp


def test_generated_header_is_current():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fq_unsat.cuh")
    assert open(here).read() == U.generate()


# ---- replay of lazy_madd (csrc/fq_lazy.cuh) on the emulated routines -------------------------------
def _wire(v):
    m = P.fq_to_mont(v)
    return [(m >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


def _mul(a, b):
    return U.run("fqu_mul", a, b)


def _set_affine(pt):
    one = U.limbs(U.ONE_INT)
    return {"x": _mul(U.run("fqu_from_wire", _wire(pt[0])), one), "y": _mul(U.run("fqu_from_wire", _wire(pt[1])), one),
            "zz": one, "zzz": one}


def _madd(acc, pt):
    u2 = _mul(U.run("fqu_from_wire", _wire(pt[0])), acc["zz"])
    s2 = _mul(U.run("fqu_from_wire", _wire(pt[1])), acc["zzz"])
    p = U.run("fqu_sub_k8", u2, acc["x"])
    r = U.run("fqu_sub_k4", s2, acc["y"])
    pp = U.run("fqu_sqr", p)
    z, e = U.run("fqu_zero_test", pp)
    assert not (z == 0 or e == 0)
    ppp = _mul(p, pp)
    q = _mul(acc["x"], pp)
    x3 = U.run("fqu_sub2_k4", U.run("fqu_sub_k2", U.run("fqu_sqr", r), ppp), q)
    y3 = U.run("fqu_sub_k2", _mul(r, U.run("fqu_sub_k8", q, x3)), _mul(acc["y"], ppp))
    out = {"x": x3, "y": y3, "zz": _mul(acc["zz"], pp), "zzz": _mul(acc["zzz"], ppp)}
    for k, bound in (("x", 7.2), ("y", 3.2), ("zz", 1.2), ("zzz", 1.2)):
        assert all(l <= U.MASK for l in out[k]) and U.value(out[k]) < bound * PM, k
    return out


def _to_affine(acc):
    to_wire = U.limbs(U.TO_WIRE)
    c = {}
    for k in acc:
        w = U.run("fqu_pack_canonical", _mul(acc[k], to_wire))
        c[k] = P.fq_from_mont(sum(x << (32 * i) for i, x in enumerate(w)))
    return (c["x"] * pow(c["zz"], -1, PM) % PM, c["y"] * pow(c["zzz"], -1, PM) % PM)


def test_lazy_mixed_addition_chain_matches_affine_addition():
    rng = P.SplitMix64(31)
    pts = [P.g1_mul(rng.fr(), P.G1_GEN) for _ in range(12)]
    acc = _set_affine(pts[0])
    ref = pts[0]
    for pt in pts[1:]:
        acc = _madd(acc, pt)
        ref = P.g1_add(ref, pt)
        assert _to_affine(acc) == ref
    # detection of P == Q and P == -Q through the square of the x-difference
    for other, same_y in ((ref, True), (P.g1_neg(ref), False)):
        u2 = _mul(U.run("fqu_from_wire", _wire(other[0])), acc["zz"])
        s2 = _mul(U.run("fqu_from_wire", _wire(other[1])), acc["zzz"])
        pp = U.run("fqu_sqr", U.run("fqu_sub_k8", u2, acc["x"]))
        z, e = U.run("fqu_zero_test", pp)
        assert z == 0 or e == 0
        rr = U.run("fqu_sqr", U.run("fqu_sub_k4", s2, acc["y"]))
        z, e = U.run("fqu_zero_test", rr)
        assert (z == 0 or e == 0) == same_y
