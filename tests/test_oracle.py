"""Pins the C oracle (oracle/oracle.c): against the big-integer oracle, the committed golden
vectors and every known answer the reference's own tests hold for the path."""
import hashlib

import numpy as np
import pytest

from oracle import pyref as P
from vectors import ORDERINGS, fr_array, fr_ints, load_golden, msm_case_inputs, ntt_case_input

GOLD = load_golden()


def mont(vals):
    return fr_array([P.fr_to_mont(v) for v in vals])


def test_constants_match_spec_literals(oracle):
    """test_known_answer_vectors.cu:60-200 (moduli, R, R^2, -m^-1, omega, generators)."""
    c = oracle.constants()
    assert P.from_limbs(c[0:6]) == 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    assert P.from_limbs(c[18:22]) == 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    assert P.from_limbs(c[6:12]) == P.FQ_R == int(GOLD["field"]["fq_one_mont"], 16)
    assert P.from_limbs(c[22:26]) == P.FR_R == int(GOLD["field"]["fr_one_mont"], 16)
    assert P.from_limbs(c[12:18]) == P.FQ_R2 and P.from_limbs(c[26:30]) == P.FR_R2
    assert int(c[30]) == 0x89F3FFFCFFFCFFFD and int(c[31]) == 0xFFFFFFFEFFFFFFFF
    assert P.from_limbs(c[32:36]) == int(GOLD["field"]["fr_root_of_unity_mont"], 16)
    assert pow(P.FR_ROOT_OF_UNITY, 1 << 32, P.R_MOD) == 1 and pow(P.FR_ROOT_OF_UNITY, 1 << 31, P.R_MOD) == P.R_MOD - 1
    assert oracle.generator(1).tobytes().hex() == GOLD["field"]["g1_gen_mont"]
    assert oracle.generator(2).tobytes().hex() == GOLD["field"]["g2_gen_mont"]
    assert oracle.on_curve(1, oracle.generator(1)) and oracle.on_curve(2, oracle.generator(2))


def test_field_known_answers_and_axioms(oracle):
    """1*1=1, 0*1=0 (test_known_answer_vectors.cu:221-236); 2*3=6, a*a^-1=1, axioms (test_field_properties.cu)."""
    one = P.to_limbs(P.FR_R, 4)
    assert P.from_limbs(oracle.fr_binop("mul", one, one)) == P.FR_R
    assert P.from_limbs(oracle.fr_binop("mul", [0] * 4, one)) == 0
    m = lambda v: P.to_limbs(P.fr_to_mont(v), 4)
    assert P.from_limbs(oracle.fr_binop("mul", m(2), m(3))) == P.fr_to_mont(6)
    rng = P.SplitMix64(2024)
    for _ in range(100):
        a, b = rng.fr(), rng.fr()
        assert P.from_limbs(oracle.fr_binop("mul", m(a), m(b))) == P.fr_to_mont(a * b % P.R_MOD)
        assert P.from_limbs(oracle.fr_binop("add", m(a), m(b))) == P.fr_to_mont((a + b) % P.R_MOD)
        assert P.from_limbs(oracle.fr_binop("sub", m(a), m(b))) == P.fr_to_mont((a - b) % P.R_MOD)
        ai = oracle.unop("fr_inv", m(a), 4)
        assert P.from_limbs(oracle.fr_binop("mul", m(a), ai)) == P.FR_R
        x, y = (a * b) % P.P_MOD, (a + 7 * b) % P.P_MOD
        q = lambda v: P.to_limbs(P.fq_to_mont(v), 6)
        assert P.from_limbs(oracle.fq_binop("mul", q(x), q(y))) == P.fq_to_mont(x * y % P.P_MOD)
        assert P.from_limbs(oracle.fq_binop("sub", q(x), q(y))) == P.fq_to_mont((x - y) % P.P_MOD)
    assert P.from_limbs(oracle.unop("fr_inv", [0] * 4, 4)) == 0          # inv(0) = 0, field.cuh:750-900
    assert P.from_limbs(oracle.unop("fr_from_mont", m(12345), 4)) == 12345
    assert P.from_limbs(oracle.unop("fr_to_mont", P.to_limbs(12345, 4), 4)) == P.fr_to_mont(12345)


@pytest.mark.parametrize("case", GOLD["g1_msm"], ids=[c["name"] for c in GOLD["g1_msm"]])
def test_g1_msm_golden(oracle, case):
    sc, bases = msm_case_inputs(case, "g1")
    assert oracle.msm(1, sc, bases).tobytes().hex() == case["result"]
    # Montgomery-form scalars and an explicit window size give the same bytes
    assert oracle.msm(1, mont(fr_ints(sc)), bases, scalars_mont=True, c=5).tobytes().hex() == case["result"]


@pytest.mark.parametrize("case", GOLD["g2_msm"], ids=[c["name"] for c in GOLD["g2_msm"]])
def test_g2_msm_golden(oracle, case):
    sc, bases = msm_case_inputs(case, "g2")
    assert oracle.msm(2, sc, bases).tobytes().hex() == case["result"]


def test_group_law_identities(oracle):
    """2P = P+P, O+P = P, P + (-P) = O   (test_curve_operations.cu / test_point_ops.cu)."""
    g = oracle.generator(1)
    two = oracle.msm(1, fr_array([2]), g.reshape(1, 12))
    pp = oracle.msm(1, fr_array([1, 1]), np.stack([g, g]))
    assert two.tobytes() == pp.tobytes() == P.g1_result_std_bytes(P.g1_add(P.G1_GEN, P.G1_GEN))
    inf = np.zeros(12, dtype=np.uint64)
    assert oracle.msm(1, fr_array([9, 1]), np.stack([inf, g])).tobytes() == P.g1_result_std_bytes(P.G1_GEN)
    neg = np.frombuffer(P.g1_affine_mont_bytes(P.g1_neg(P.G1_GEN)), dtype=np.uint64)
    assert oracle.msm(1, fr_array([3, 3]), np.stack([g, neg])).tobytes() == P.g1_result_std_bytes(None)


def test_series_points_and_dlog_msm(oracle):
    rng = P.SplitMix64(77)
    k0, d = rng.fr(), rng.fr()
    n = 2048
    pts = oracle.gen_series(1, P.to_limbs(k0, 4), P.to_limbs(d, 4), n)
    for i in (0, 1, 1023, 1024, n - 1):
        assert pts[i].tobytes() == P.g1_affine_mont_bytes(P.g1_mul((k0 + i * d) % P.R_MOD, P.G1_GEN))
        assert oracle.on_curve(1, pts[i])
    sc = oracle.random_fr(5, n)
    exp = sum(s * ((k0 + i * d) % P.R_MOD) for i, s in enumerate(fr_ints(sc))) % P.R_MOD
    assert oracle.msm(1, sc, pts).tobytes() == P.g1_result_std_bytes(P.g1_mul(exp, P.G1_GEN))
    kk = fr_array([(k0 + i * d) % P.R_MOD for i in range(n)])
    assert P.from_limbs(oracle.fr_dot(sc, kk)) == exp
    pts2 = oracle.gen_series(2, P.to_limbs(k0, 4), P.to_limbs(d, 4), 64)
    assert pts2[63].tobytes() == P.g2_affine_mont_bytes(P.g2_mul((k0 + 63 * d) % P.R_MOD, P.G2_GEN))
    exp2 = sum(s * ((k0 + i * d) % P.R_MOD) for i, s in enumerate(fr_ints(sc[:64]))) % P.R_MOD
    assert oracle.msm(2, sc[:64], pts2).tobytes() == P.g2_result_std_bytes(P.g2_mul(exp2, P.G2_GEN))


def test_random_fr_matches_pyref(oracle):
    a = oracle.random_fr(0xB12381, 50)
    rng = P.SplitMix64(0xB12381)
    assert fr_ints(a) == [rng.fr() for _ in range(50)]


def run_ntt_case(oracle, case):
    vec = ntt_case_input(case)
    o = case["ordering"]
    nat = P.apply_ordering(vec, o, "in")
    a = mont(nat)
    if case["coset"]:
        out = oracle.coset_ntt(a, mont([case["coset"]])[0], inverse=case["inverse"])
    else:
        out = oracle.ntt(a, inverse=case["inverse"])
    if o[1] == "R":
        out = oracle.bit_reverse(out)
    return out


@pytest.mark.parametrize("case", GOLD["ntt"], ids=[c["name"] for c in GOLD["ntt"]])
def test_ntt_golden(oracle, case):
    out = run_ntt_case(oracle, case)
    assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"]
    assert [hex(v) for v in fr_ints(out[: len(case["head"])])] == case["head"]


def test_ntt_properties(oracle):
    """round trip, zeros, linearity, convolution theorem, delta (test_ntt_security.cu:993-1013)."""
    n = 256
    a, b = oracle.random_fr(1, n), oracle.random_fr(2, n)
    A, B = oracle.ntt(a), oracle.ntt(b)
    assert (oracle.ntt(A, inverse=True) == a).all()
    assert not oracle.ntt(np.zeros((n, 4), dtype=np.uint64)).any()
    assert (oracle.ntt(oracle.vecop(0, a, b)) == oracle.vecop(0, A, B)).all()
    # cyclic convolution via pointwise product
    ai, bi = [P.fr_from_mont(v) for v in fr_ints(a)], [P.fr_from_mont(v) for v in fr_ints(b)]
    conv = [sum(ai[j] * bi[(i - j) % n] for j in range(n)) % P.R_MOD for i in (0, 1, 100, n - 1)]
    got = oracle.ntt(oracle.vecop(2, A, B), inverse=True)
    assert [P.fr_from_mont(v) for v in fr_ints(got[[0, 1, 100, n - 1]])] == conv
    assert P.from_limbs(oracle.omega(16)) == P.fr_to_mont(pow(P.FR_ROOT_OF_UNITY, 1 << 16, P.R_MOD))  # ntt_fft_comparison.rs:133-173


def test_vecops_golden(oracle):
    v = GOLD["vecops"][0]
    a = fr_array([int(h, 16) for h in v["a"]])
    b = fr_array([int(h, 16) for h in v["b"]])
    for op, name in ((0, "add"), (1, "sub"), (2, "mul")):
        assert [hex(x) for x in fr_ints(oracle.vecop(op, a, b))] == v[name]
    assert [hex(x) for x in fr_ints(oracle.vecop(2, a[0], b, a_scalar=True))] == v["scalar_mul"]
    assert [hex(x) for x in fr_ints(oracle.vecop(0, a[0], b, a_scalar=True))] == v["scalar_add"]


def test_poly_eval_pins_ntt_outputs(oracle):
    """oracle.poly_eval (Horner, shares no code with orc_ntt) against big integers and against single NTT outputs:
    y[i] = A(omega^i).  The GPU tests use it for spot checks at 2^24."""
    n = 1 << 10
    a = oracle.random_fr(99, n)
    ai = [P.fr_from_mont(v) for v in fr_ints(a)]
    y = oracle.ntt(a)
    w = P.fr_omega(10)
    for i in (0, 1, 517, n - 1):
        z = pow(w, i, P.R_MOD)
        got = oracle.poly_eval(a, mont([z])[0])
        assert P.fr_from_mont(P.from_limbs(got)) == sum(c * pow(z, j, P.R_MOD) for j, c in enumerate(ai)) % P.R_MOD
        assert (got == y[i]).all()
    # chunked path (n > 2^16) against the transform
    n = 1 << 18
    a = oracle.random_fr(100, n)
    y = oracle.ntt(a)
    w = P.fr_omega(18)
    for i in (3, n // 2 + 1, n - 1):
        assert (oracle.poly_eval(a, mont([pow(w, i, P.R_MOD)])[0]) == y[i]).all()


def test_endomorphism_constants():
    """The constants csrc/gen/gen_field.py writes for csrc/glv.cuh equal the ones oracle/pyref.py derives on its own;
    beta and lambda also equal the literals the reference carries (bls12-381/src/curve/point_ops.cu:120-133:
    GLV_BETA Montgomery, GLV_LAMBDA), and phi / psi act as [lambda] / [z] on the generators."""
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "midnight_bls12_381_cuda_b200",
                            "csrc", "field_consts.h")).read()

    def limbs(name):
        body = re.search(r"#define " + name + r" (.*)", hdr).group(1).split("/*")[0]
        return [int(x, 16) for x in re.findall(r"0x([0-9a-f]+)ull", body)]

    assert P.from_limbs(limbs("GLV_BETA_MONT_INIT")) == P.fq_to_mont(P.GLV_BETA)
    assert limbs("GLV_BETA_MONT_INIT") == [0xcd03c9e48671f071, 0x5dab22461fcda5d2, 0x587042afd3851b95, 0x8eb60ebe01bacb9e,
                                           0x03f97d6e83d050d2, 0x18f0206554638741]
    assert P.from_limbs(limbs("GLV_LAMBDA_INIT")) == P.GLV_LAMBDA == 0xac45a4010001a40200000000ffffffff
    assert P.from_limbs(limbs("GLV_RECIP_INIT")) == (1 << 256) // P.GLV_LAMBDA
    assert limbs("BLS_Z_ABS")[0] == -P.BLS_X
    cx, cy = limbs("PSI_CX_MONT_INIT"), limbs("PSI_CY_MONT_INIT")
    assert (P.from_limbs(cx[:6]), P.from_limbs(cx[6:])) == (P.fq_to_mont(P.PSI_CX.c0), P.fq_to_mont(P.PSI_CX.c1))
    assert (P.from_limbs(cy[:6]), P.from_limbs(cy[6:])) == (P.fq_to_mont(P.PSI_CY.c0), P.fq_to_mont(P.PSI_CY.c1))
    assert pow(P.GLV_BETA, 3, P.P_MOD) == 1 and P.GLV_BETA != 1
    k = 0x1234567
    pt = P.g1_mul(k, P.G1_GEN)
    assert (pt[0] * P.GLV_BETA % P.P_MOD, pt[1]) == P.g1_mul(P.GLV_LAMBDA, pt)
    q = P.g2_mul(k, P.G2_GEN)
    zq = P.g2_mul(P.BLS_X % P.R_MOD, q)
    psi = P.g2_psi(q)
    assert psi[0] == zq[0] and psi[1] == zq[1] and P.g2_on_curve(psi)
    # ground truth of the membership tests: random curve points are not members, generator multiples are
    assert P.g1_on_curve(P.g1_curve_point(3)) and not P.g1_in_subgroup(P.g1_curve_point(3)) and P.g1_in_subgroup(pt)
    assert P.g2_on_curve(P.g2_curve_point(5)) and not P.g2_in_subgroup(P.g2_curve_point(5)) and P.g2_in_subgroup(q)
