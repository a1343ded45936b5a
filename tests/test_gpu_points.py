"""Point-format conversions and on-curve validation on the GPU, through the C ABI with
the reference's entry-point names (bls12-381/src/curve/point_ops.cu:759-1000), against Python big integers."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyref as P

pytestmark = pytest.mark.gpu


def fq_m(v):
    return P.fq_bytes(P.fq_to_mont(v % P.P_MOD))


def g1_jac_bytes(pt, z):
    """(x z^2, y z^3, z) Montgomery; infinity = (0, R, 0) (point.cuh:469-486)."""
    if pt is None:
        return fq_m(0) + fq_m(1) + fq_m(0)
    x, y = pt
    return fq_m(x * z * z) + fq_m(y * z * z * z) + fq_m(z)


def fq2_m(v):
    return fq_m(v.c0) + fq_m(v.c1)


def g2_jac_bytes(pt, z):
    if pt is None:
        return fq2_m(P.Fq2(0, 0)) + fq2_m(P.Fq2(1, 0)) + fq2_m(P.Fq2(0, 0))
    x, y = pt
    return fq2_m(x * z * z) + fq2_m(y * z * z * z) + fq2_m(z)


def call(b381, name, inp, n, out, on_device=False):
    lib = b381.lib()
    cfg = lib.b381_default_vecops_config()
    cfg.is_a_on_device = cfg.is_result_on_device = on_device
    return getattr(lib, name)(b381.ptr(inp), n, C.byref(cfg), b381.ptr(out))


@pytest.mark.parametrize("n", [1, 5, 33, 1000])
def test_g1_conversions(cuda, b381, n):
    rng = P.SplitMix64(900 + n)
    pts = [P.g1_mul(rng.fr(), P.G1_GEN) for _ in range(min(n, 40))]
    pts = (pts * (n // len(pts) + 1))[:n]
    zs = [rng.fr() % P.P_MOD or 1 for _ in range(n)]
    if n > 3:
        pts[2] = None
        zs[0] = 1
    jac = np.frombuffer(b"".join(g1_jac_bytes(p, z) for p, z in zip(pts, zs)), dtype=np.uint64).copy()
    aff_exp = b"".join(P.g1_affine_mont_bytes(p) for p in pts)
    out = np.zeros(12 * n, dtype=np.uint64)
    assert call(b381, "bls12_381_g1_projective_to_affine", jac, n, out) == 0
    assert out.tobytes() == aff_exp
    # device-resident in/out
    d_in, d_out = cuda.from_numpy(jac.view(np.int64)).cuda(), cuda.zeros(12 * n, dtype=cuda.int64, device="cuda")
    assert call(b381, "bls12_381_g1_projective_to_affine", d_in, n, d_out, on_device=True) == 0
    assert d_out.cpu().numpy().tobytes() == aff_exp
    # affine -> Jacobian is (x, y, R) / (0, R, 0), and converting back is the identity
    aff = np.frombuffer(aff_exp, dtype=np.uint64).copy()
    j2 = np.zeros(18 * n, dtype=np.uint64)
    assert call(b381, "bls12_381_g1_affine_to_projective", aff, n, j2) == 0
    assert j2.tobytes() == b"".join(g1_jac_bytes(p, 1) for p in pts)
    back = np.zeros(12 * n, dtype=np.uint64)
    assert call(b381, "bls12_381_g1_projective_to_affine", j2, n, back) == 0
    assert back.tobytes() == aff_exp


def test_g2_conversions(cuda, b381):
    rng = P.SplitMix64(77)
    n = 37
    pts = [P.g2_mul(rng.fr(), P.G2_GEN) for _ in range(8)]
    pts = (pts * 5)[:n]
    pts[4] = None
    zs = [P.Fq2(rng.fr(), rng.fr()) for _ in range(n)]
    jac = np.frombuffer(b"".join(g2_jac_bytes(p, z) for p, z in zip(pts, zs)), dtype=np.uint64).copy()
    aff_exp = b"".join(P.g2_affine_mont_bytes(p) for p in pts)
    out = np.zeros(24 * n, dtype=np.uint64)
    assert call(b381, "bls12_381_g2_projective_to_affine", jac, n, out) == 0
    assert out.tobytes() == aff_exp
    j2 = np.zeros(36 * n, dtype=np.uint64)
    assert call(b381, "bls12_381_g2_affine_to_projective", np.frombuffer(aff_exp, dtype=np.uint64).copy(), n, j2) == 0
    assert j2.tobytes() == b"".join(g2_jac_bytes(p, P.Fq2(1, 0)) for p in pts)


def test_on_curve_and_argument_errors(cuda, b381):
    rng = P.SplitMix64(5)
    n = 64
    pts = [P.g1_mul(rng.fr(), P.G1_GEN) for _ in range(n)]
    pts[3] = None
    raw = bytearray(b"".join(P.g1_affine_mont_bytes(p) for p in pts))
    bad = {7, 20, 63}
    for i in bad:                      # corrupt y: off the curve
        raw[96 * i + 48] ^= 1
    flags = np.zeros(n, dtype=np.uint8)
    assert call(b381, "b381_g1_is_on_curve", np.frombuffer(bytes(raw), dtype=np.uint64).copy(), n, flags) == 0
    assert [i for i in range(n) if not flags[i]] == sorted(bad)
    pts2 = [P.g2_mul(rng.fr(), P.G2_GEN) for _ in range(9)] + [None]
    raw2 = bytearray(b"".join(P.g2_affine_mont_bytes(p) for p in pts2))
    raw2[192 * 2 + 5] ^= 4
    f2 = np.zeros(10, dtype=np.uint8)
    assert call(b381, "b381_g2_is_on_curve", np.frombuffer(bytes(raw2), dtype=np.uint64).copy(), 10, f2) == 0
    assert list(f2) == [1, 1, 0, 1, 1, 1, 1, 1, 1, 1]
    # the reference's argument checks (point_ops.cu:766-775): null pointers and size <= 0 -> INVALID_ARGUMENT
    lib = b381.lib()
    cfg = lib.b381_default_vecops_config()
    assert lib.bls12_381_g1_projective_to_affine(None, 4, C.byref(cfg), b381.ptr(flags)) == 11
    assert lib.bls12_381_g1_projective_to_affine(b381.ptr(flags), 0, C.byref(cfg), b381.ptr(flags)) == 11


# ---------------------------------------------------------------- endomorphisms: GLV scalar multiplication, subgroup checks
def test_g1_scalar_mul_glv_and_plain(cuda, b381):
    """bls12_381_g1_scalar_mul_glv / bls12_381_g1_scalar_mul (point_ops.cu:1019-1268): out[i] = k_i * P_i, canonical
    scalars, Montgomery affine bases; both paths against big-integer k * P, byte for byte in the normalised Jacobian
    form, host and device residency, including the scalars at the edges of the k = k1 + k2 * lambda split."""
    import vectors_points as V
    rng = P.SplitMix64(4242)
    ks = V.glv_scalars() + [rng.fr() for _ in range(34)]
    n = len(ks)
    base_pts = [P.g1_mul(rng.fr(), P.G1_GEN) for _ in range(16)]
    pts = [base_pts[i % 16] for i in range(n)]
    pts[5] = None
    bases = np.frombuffer(b"".join(P.g1_affine_mont_bytes(p) for p in pts), dtype=np.uint64).copy()
    sc = np.frombuffer(b"".join(P.fr_bytes(k) for k in ks), dtype=np.uint64).copy()
    exp = b"".join(g1_jac_bytes(P.g1_mul(k, p), 1) for k, p in zip(ks, pts))
    lib = b381.lib()
    for name in ("bls12_381_g1_scalar_mul_glv", "bls12_381_g1_scalar_mul"):
        cfg = lib.b381_default_vecops_config()
        out = np.zeros(18 * n, dtype=np.uint64)
        assert getattr(lib, name)(b381.ptr(bases), b381.ptr(sc), n, C.byref(cfg), b381.ptr(out)) == 0
        assert out.tobytes() == exp, name
        cfg.is_a_on_device = cfg.is_b_on_device = cfg.is_result_on_device = True
        d_b, d_s = cuda.from_numpy(bases.view(np.int64)).cuda(), cuda.from_numpy(sc.view(np.int64)).cuda()
        d_o = cuda.zeros(18 * n, dtype=cuda.int64, device="cuda")
        assert getattr(lib, name)(b381.ptr(d_b), b381.ptr(d_s), n, C.byref(cfg), b381.ptr(d_o)) == 0
        assert d_o.cpu().numpy().tobytes() == exp, name
        assert getattr(lib, name)(None, b381.ptr(sc), n, C.byref(cfg), b381.ptr(out)) == 11
        assert getattr(lib, name)(b381.ptr(bases), b381.ptr(sc), 0, C.byref(cfg), b381.ptr(out)) == 11


def test_g1_scalar_mul_glv_matches_msm(cuda, b381):
    """sum of the GLV products == the MSM over the same inputs (two independent code paths), 2^12 points"""
    n = 1 << 12
    lib = b381.lib()
    g = np.frombuffer(P.g1_affine_mont_bytes(P.G1_GEN), dtype=np.uint64).copy()
    bases = cuda.empty((n, 12), dtype=cuda.int64, device="cuda")
    assert lib.b381_g1_point_series(b381.ptr(g), b381.ptr(g), C.c_uint64(n), b381.ptr(bases), None) == 0
    sc_int = fr_ints_local(n, 99)
    sc = cuda.from_numpy(np.frombuffer(b"".join(P.fr_bytes(k) for k in sc_int), dtype=np.uint64).copy().view(np.int64)).cuda()
    cfg = lib.b381_default_vecops_config()
    cfg.is_a_on_device = cfg.is_b_on_device = cfg.is_result_on_device = True
    prods = cuda.zeros((n, 18), dtype=cuda.int64, device="cuda")
    assert lib.bls12_381_g1_scalar_mul_glv(b381.ptr(bases), b381.ptr(sc), n, C.byref(cfg), b381.ptr(prods)) == 0
    # spot check 8 products against big integers, and the sum against the discrete log
    host = prods.cpu().numpy().view(np.uint64)
    for i in (0, 1, 77, 500, 1023, 2048, 4000, n - 1):
        assert host[i].tobytes() == g1_jac_bytes(P.g1_mul(sc_int[i] * (i + 1), P.G1_GEN), 1)
    mcfg = lib.b381_default_msm_config()
    mcfg.are_scalars_on_device = mcfg.are_points_on_device = True
    mcfg.are_points_montgomery_form = True
    res = np.zeros(18, dtype=np.uint64)
    assert lib.b381_g1_msm(b381.ptr(sc), b381.ptr(bases), n, C.byref(mcfg), b381.ptr(res)) == 0
    dl = sum(k * (i + 1) for i, k in enumerate(sc_int)) % P.R_MOD
    assert res.tobytes() == P.g1_result_std_bytes(P.g1_mul(dl, P.G1_GEN))


def fr_ints_local(n, seed):
    rng = P.SplitMix64(seed)
    return [rng.fr() for _ in range(n)]


def test_subgroup_checks(cuda, b381):
    """b381_g1_is_in_subgroup / b381_g2_is_in_subgroup against [r]P == O computed with big integers: members,
    infinity, random curve points, pure cofactor-subgroup points (one of order 3), member + non-member sums."""
    import vectors_points as V
    lib = b381.lib()
    for name, cases, enc, words in (("b381_g1_is_in_subgroup", V.g1_membership_cases(), P.g1_affine_mont_bytes, 12),
                                    ("b381_g2_is_in_subgroup", V.g2_membership_cases(), P.g2_affine_mont_bytes, 24)):
        cases = cases * 3                         # more than one warp's worth of mixed outcomes
        n = len(cases)
        raw = np.frombuffer(b"".join(enc(p) for p, _ in cases), dtype=np.uint64).copy()
        flags = np.full(n, 7, dtype=np.uint8)
        assert call(b381, name, raw, n, flags) == 0
        assert list(flags) == [int(m) for _, m in cases], name
        d_in = cuda.from_numpy(raw.view(np.int64)).cuda()
        d_f = cuda.zeros(n, dtype=cuda.uint8, device="cuda")
        assert call(b381, name, d_in, n, d_f, on_device=True) == 0
        assert list(d_f.cpu().numpy()) == [int(m) for _, m in cases], name
        cfg = lib.b381_default_vecops_config()
        assert getattr(lib, name)(None, n, C.byref(cfg), b381.ptr(flags)) == 11
        assert getattr(lib, name)(b381.ptr(raw), -1, C.byref(cfg), b381.ptr(flags)) == 11


def test_points_host_layer(cuda, b381):
    """midnight_bls12_381_cuda_b200.points: the numpy-level wrappers over the same entry points"""
    import midnight_bls12_381_cuda_b200 as M
    import vectors_points as V
    rng = P.SplitMix64(11)
    pts = [P.g1_mul(rng.fr(), P.G1_GEN) for _ in range(9)] + [None]
    aff = np.frombuffer(b"".join(P.g1_affine_mont_bytes(p) for p in pts), dtype=np.uint64).reshape(-1, 12)
    jac = M.points.g1_affine_to_projective(aff)
    assert (M.points.g1_projective_to_affine(jac) == aff).all()
    assert M.points.g1_is_on_curve(aff).all() and M.points.g1_is_in_subgroup(aff).all()
    M.points.validate_g1_bases(aff)
    cases = V.g1_membership_cases()
    raw = np.frombuffer(b"".join(P.g1_affine_mont_bytes(p) for p, _ in cases), dtype=np.uint64).reshape(-1, 12)
    assert list(M.points.g1_is_in_subgroup(raw)) == [m for _, m in cases]
    with pytest.raises(M.points.PointError):
        M.points.validate_g1_bases(raw)
    ks = [rng.fr() for _ in pts]
    sc = np.frombuffer(b"".join(P.fr_bytes(k) for k in ks), dtype=np.uint64).reshape(-1, 4)
    exp = b"".join(g1_jac_bytes(P.g1_mul(k, p), 1) for k, p in zip(ks, pts))
    assert M.points.g1_scalar_mul(aff, sc).tobytes() == exp == M.points.g1_scalar_mul(aff, sc, glv=False).tobytes()
