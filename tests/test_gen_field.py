"""The generated inline-PTX field routines, executed by the Python PTX interpreter
(csrc/gen/ptxir.py) against big integers -- no GPU needed.  Mirrors the reference's
test_field_properties.cu (field axioms on random Montgomery values) and the Fr KATs
1*1=1, 0*1=0, 2*3=6 (tests/test_known_answer_vectors.cu:221-236)."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "midnight_bls12_381_cuda_b200", "csrc", "gen"))
import gen_field as G  # noqa: E402


@pytest.mark.parametrize("f", [G.FQ, G.FR], ids=["fq", "fr"])
def test_ptx_routines_match_bigint(f):
    rnd = random.Random(1234)
    rinv = pow(f.R, -1, f.m)
    blk = {op: G.build(f, op) for op in G.OPS}
    edge = [0, 1, 2, f.m - 1, f.m - 2, f.R, f.R2, (1 << (32 * f.n - 1)) % f.m, f.m >> 1]
    vals = edge + [rnd.randrange(f.m) for _ in range(120)]
    for i, a in enumerate(vals):
        for b in (vals[(7 * i + 3) % len(vals)], a, vals[(13 * i + 1) % len(vals)]):
            assert G.run_block(blk["mul"], f, a, b) == a * b * rinv % f.m
            assert G.run_block(blk["add"], f, a, b) == (a + b) % f.m
            assert G.run_block(blk["sub"], f, a, b) == (a - b) % f.m
        assert G.run_block(blk["sqr"], f, a) == a * a * rinv % f.m
        assert G.run_block(blk["neg"], f, a) == (-a) % f.m
        assert G.run_block(blk["dbl"], f, a) == 2 * a % f.m


@pytest.mark.parametrize("f", [G.FQ, G.FR], ids=["fq", "fr"])
def test_mul_with_degenerate_limbs(f):
    """operands whose 32-bit limbs are 0 or 0xffffffff: every reduction row of Fr's specialised Montgomery step
    (m_i = -E[0], the products by the modulus limbs 1 and 0xffffffff replaced by adds) sees E[0] = 0 and
    E[0] = 0xffffffff, and the carries out of the replaced pairs are exercised."""
    rnd = random.Random(99)
    rinv = pow(f.R, -1, f.m)
    mul = G.build(f, "mul")
    vals = [1 << 32, 1 << 64, (1 << 96) - (1 << 32), f.m - (1 << 32), 0xFFFFFFFF, 0xFFFFFFFF << 32, (1 << (32 * f.n - 2)) - 1]
    for _ in range(150):
        limbs = [rnd.choice((0, 0, 0xFFFFFFFF, 0xFFFFFFFF, 1, rnd.getrandbits(32))) for _ in range(f.n)]
        vals.append(sum(l << (32 * k) for k, l in enumerate(limbs)) % f.m)
    sqr = G.build(f, "sqr")
    for i, a in enumerate(vals):
        for b in (vals[(5 * i + 1) % len(vals)], a, rnd.randrange(f.m)):
            assert G.run_block(mul, f, a, b) == a * b * rinv % f.m, (hex(a), hex(b))
        # Fq's one-sided squaring (mont_sqr_body): the doubled limbs and the carry-only pairs see all-0 / all-1 limbs
        assert G.run_block(sqr, f, a) == a * a * rinv % f.m, hex(a)


def test_fq_squaring_takes_each_cross_product_once():
    """78 + 144 wide multiply-adds (+ 12 m_i) instead of 288: the IMAD-pipe saving the squaring exists for."""
    def wide(blk):
        c = blk.count()
        return sum(v for k, v in c.items() if ".hi" in k)          # one IMAD.WIDE per lo/hi pair
    assert wide(G.build(G.FQ, "mul")) == 288 and wide(G.build(G.FQ, "sqr")) == 222
    assert wide(G.build(G.FR, "sqr")) == wide(G.build(G.FR, "mul"))  # Fr: accumulator bound fails, stays mul(a, a)


def test_fr_known_answers():
    f = G.FR
    mul = G.build(f, "mul")
    one = f.R
    assert G.run_block(mul, f, one, one) == one            # 1*1 = 1
    assert G.run_block(mul, f, 0, one) == 0                # 0*1 = 0
    two, three, six = 2 * f.R % f.m, 3 * f.R % f.m, 6 * f.R % f.m
    assert G.run_block(mul, f, two, three) == six          # 2*3 = 6
    a = 0x1234567890ABCDEF1234567890ABCDEF % f.m
    am, ainv = a * f.R % f.m, pow(a, -1, f.m) * f.R % f.m
    assert G.run_block(mul, f, am, ainv) == one            # a * a^-1 = 1


def test_generated_header_is_current():
    """field_ptx.cuh on disk is what the generator emits (nobody hand-edited it)."""
    here = os.path.join(os.path.dirname(__file__), "..", "midnight_bls12_381_cuda_b200", "csrc", "field_ptx.cuh")
    assert open(here).read() == G.generate()
