"""G1/G2 MSM parity on the GPU, through the C ABI (via the core/msm.rs-shaped Python layer):
golden vectors, oracle.c on the same seeded inputs, edge cases of the reference's tests
(test_msm_security.cu:908-941, core/msm.rs:1653-2111) and size-independent checks at full size."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyref as P
from vectors import fr_array, fr_ints, load_golden, msm_case_inputs

pytestmark = pytest.mark.gpu
GOLD = load_golden()


def mont(sc_int_arr):
    return fr_array([P.fr_to_mont(v) for v in fr_ints(sc_int_arr)])


def raw_msm(b381, group, scalars, bases, n, **kw):
    """direct C-ABI call with explicit flags; host buffers unless told otherwise."""
    lib = b381.lib()
    cfg = lib.b381_default_msm_config()
    cfg.are_points_montgomery_form = kw.get("points_mont", True)
    cfg.are_scalars_montgomery_form = kw.get("scalars_mont", False)
    cfg.c = kw.get("c", 0)
    cfg.bitsize = kw.get("bitsize", 0)
    cfg.batch_size = kw.get("batch", 1)
    cfg.are_points_shared_in_batch = kw.get("shared", True)
    cfg.precompute_factor = kw.get("factor", 1)
    cfg.are_scalars_on_device = kw.get("scalars_on_device", False)
    cfg.are_points_on_device = kw.get("points_on_device", False)
    k = 1 if group == "g1" else 2
    res = np.zeros((cfg.batch_size, 18 * k), dtype=np.uint64)
    fn = lib.b381_g1_msm if k == 1 else lib.b381_g2_msm
    code = fn(b381.ptr(scalars), b381.ptr(bases), n, C.byref(cfg), b381.ptr(res))
    assert code == 0, b381.ERROR_NAMES.get(code, code)
    return res


@pytest.mark.parametrize("case", GOLD["g1_msm"], ids=[c["name"] for c in GOLD["g1_msm"]])
def test_g1_golden(cuda, b381, case):
    sc, bases = msm_case_inputs(case, "g1")
    n = case["n"]
    assert raw_msm(b381, "g1", sc, bases, n).tobytes().hex() == case["result"]
    assert raw_msm(b381, "g1", mont(sc), bases, n, scalars_mont=True, c=7).tobytes().hex() == case["result"]
    # the reference's own flat test entry point: Jacobian Montgomery result (icicle_curve_api.cu:679-706)
    lib = b381.lib()
    cfg = lib.b381_default_msm_config()
    res = np.zeros(18, dtype=np.uint64)
    assert lib.bls12_381_g1_msm_cuda(b381.ptr(sc), b381.ptr(bases), n, C.byref(cfg), b381.ptr(res)) == 0
    z = P.from_limbs(res[12:18])
    exp = bytes.fromhex(case["result"])
    if z == 0:
        assert exp[96:] == bytes(48)
    else:
        assert z == P.FQ_R
        assert P.fq_from_mont(P.from_limbs(res[0:6])).to_bytes(48, "little") == exp[:48]
        assert P.fq_from_mont(P.from_limbs(res[6:12])).to_bytes(48, "little") == exp[48:96]


@pytest.mark.parametrize("case", GOLD["g2_msm"], ids=[c["name"] for c in GOLD["g2_msm"]])
def test_g2_golden(cuda, b381, case):
    sc, bases = msm_case_inputs(case, "g2")
    assert raw_msm(b381, "g2", sc, bases, case["n"]).tobytes().hex() == case["result"]
    assert raw_msm(b381, "g2", mont(sc), bases, case["n"], scalars_mont=True, c=5).tobytes().hex() == case["result"]


@pytest.mark.parametrize("logn,seed", [(10, 1), (14, 2), (16, 0xB12381_0016)])
def test_g1_vs_oracle_random(cuda, b381, oracle, logn, seed):
    """config 1 of BASELINE.json: random points/scalars, CPU path vs new kernel, bit-exact."""
    n = 1 << logn
    rng = P.SplitMix64(seed)
    k0, d = rng.fr(), rng.fr()
    bases = oracle.gen_series(1, P.to_limbs(k0, 4), P.to_limbs(d, 4), n)
    sc = oracle.random_fr(seed + 1000, n)
    exp = oracle.msm(1, sc, bases)
    assert raw_msm(b381, "g1", sc, bases, n).tobytes() == exp.tobytes()
    # through the core/msm.rs-shaped API with device-resident bases and Montgomery scalars
    import midnight_bls12_381_cuda_b200 as M
    ctx = M.GpuMsmContext()
    dev = ctx.upload_g1_bases(bases)
    got = ctx.msm_with_device_bases(mont(sc), dev)
    x, y = got
    assert x.to_bytes(48, "little") + y.to_bytes(48, "little") == exp.tobytes()[:96]
    assert ctx.msm_with_device_bases_async(mont(sc), dev).wait() == got


def test_g2_vs_oracle_random(cuda, b381, oracle):
    n = 1 << 10
    rng = P.SplitMix64(404)
    k0, d = rng.fr(), rng.fr()
    bases = oracle.gen_series(2, P.to_limbs(k0, 4), P.to_limbs(d, 4), n)
    sc = oracle.random_fr(405, n)
    assert raw_msm(b381, "g2", sc, bases, n).tobytes() == oracle.msm(2, sc, bases).tobytes()


def test_edge_cases(cuda, b381, oracle):
    g = oracle.generator(1)
    inf = np.zeros(12, dtype=np.uint64)
    # n = 0 -> identity (test_msm_security.cu "empty"), results (0,1,0)
    assert raw_msm(b381, "g1", fr_array([]), np.zeros((0, 12), dtype=np.uint64), 0).tobytes() == P.g1_result_std_bytes(None)
    # all-zero scalars, infinity bases, mixed zeros
    n = 300
    bases = np.tile(g, (n, 1))
    assert raw_msm(b381, "g1", fr_array([0] * n), bases, n).tobytes() == P.g1_result_std_bytes(None)
    sc = fr_array([i % 2 for i in range(n)])
    assert raw_msm(b381, "g1", sc, bases, n).tobytes() == P.g1_result_std_bytes(P.g1_mul(n // 2, P.G1_GEN))
    bases2 = bases.copy()
    bases2[::3] = inf
    exp = sum(i + 1 for i in range(n) if i % 3) % P.R_MOD
    assert raw_msm(b381, "g1", fr_array(list(range(1, n + 1))), bases2, n).tobytes() == P.g1_result_std_bytes(P.g1_mul(exp, P.G1_GEN))
    # P and -P with equal scalars cancel inside one bucket
    neg = np.frombuffer(P.g1_affine_mont_bytes(P.g1_neg(P.G1_GEN)), dtype=np.uint64)
    pair = np.stack([g, neg] * 8)
    assert raw_msm(b381, "g1", fr_array([12345] * 16), pair, 16).tobytes() == P.g1_result_std_bytes(None)
    # r-1 scalars, maximal digits in every window
    assert raw_msm(b381, "g1", fr_array([P.R_MOD - 1] * 5), np.tile(g, (5, 1)), 5).tobytes() == \
        P.g1_result_std_bytes(P.g1_mul(5 * (P.R_MOD - 1), P.G1_GEN))
    # standard-form (non-Montgomery) points are converted on a private copy
    std = np.frombuffer(P.fq_bytes(P.G1_X) + P.fq_bytes(P.G1_Y), dtype=np.uint64)
    assert raw_msm(b381, "g1", fr_array([7]), std.reshape(1, 12), 1, points_mont=False).tobytes() == \
        P.g1_result_std_bytes(P.g1_mul(7, P.G1_GEN))
    # invalid arguments come back as error codes, never exceptions across the ABI
    lib = b381.lib()
    cfg = lib.b381_default_msm_config()
    assert lib.b381_g1_msm(None, None, 4, C.byref(cfg), b381.ptr(np.zeros(18, dtype=np.uint64))) == 3   # INVALID_POINTER
    assert lib.b381_g1_msm(b381.ptr(sc), b381.ptr(bases), -1, C.byref(cfg), b381.ptr(np.zeros(18, dtype=np.uint64))) == 11


@pytest.mark.parametrize("levels", [0, 1, 2, 5])
def test_affine_prereduction_levels(cuda, b381, oracle, monkeypatch, levels):
    """csrc/msm_batch.cuh forced on at small sizes: every bucket halved `levels` times by affine pair sums under
    one CTA-wide batched inversion before the task/accumulate path; random inputs with long buckets, infinity
    bases, zero scalars, identical bases (every pair a doubling), P/-P pairs (every pair cancels), and G2."""
    monkeypatch.setenv("B381_MSM_LEVELS", str(levels))
    n = 5000
    bases = oracle.gen_series(1, [9, 1, 0, 0], [5, 0, 7, 0], n)
    sc = oracle.random_fr(31 + levels, n)
    sc[7] = 0
    sc[8] = sc[9]
    bases[11] = 0
    bases[12] = 0
    exp = oracle.msm(1, sc, bases).tobytes()
    for c in (4, 7, 11):
        assert raw_msm(b381, "g1", sc, bases, n, c=c).tobytes() == exp, c
    g = oracle.generator(1)
    m = 777
    assert raw_msm(b381, "g1", fr_array([5] * m), np.tile(g, (m, 1)), m, c=5).tobytes() == \
        P.g1_result_std_bytes(P.g1_mul(5 * m, P.G1_GEN))
    assert raw_msm(b381, "g1", fr_array(list(range(1, m + 1))), np.tile(g, (m, 1)), m, c=6).tobytes() == \
        P.g1_result_std_bytes(P.g1_mul(m * (m + 1) // 2, P.G1_GEN))
    neg = np.frombuffer(P.g1_affine_mont_bytes(P.g1_neg(P.G1_GEN)), dtype=np.uint64)
    pair = np.stack([g, neg] * 64)
    assert raw_msm(b381, "g1", fr_array([12345] * 128), pair, 128, c=4).tobytes() == P.g1_result_std_bytes(None)
    n2 = 600
    bases2 = oracle.gen_series(2, [3, 0, 0, 1], [1, 2, 0, 0], n2)
    sc2 = oracle.random_fr(77, n2)
    sc2[5] = sc2[6]
    bases2[6] = bases2[5]
    assert raw_msm(b381, "g2", sc2, bases2, n2, c=5).tobytes() == oracle.msm(2, sc2, bases2).tobytes()


@pytest.mark.parametrize("shift", [0, 8, 16])
def test_device_bases_alignment(cuda, b381, oracle, monkeypatch, shift):
    """The affine levels fetch points with 256-bit loads when the element is 32-byte aligned (msm_batch.cuh load_wide) and
    with plain loads otherwise: resident bases handed over at a device address that is only 8- or 16-byte aligned must
    give the same bytes (G1 and G2, levels forced on at a small size)."""
    import torch
    monkeypatch.setenv("B381_MSM_LEVELS", "3")
    for group, n, words in ((1, 3000, 12), (2, 500, 24)):
        bases = oracle.gen_series(group, [9, 1, 0, 0], [5, 0, 7, 0], n)
        sc = oracle.random_fr(400 + shift + group, n)
        exp = oracle.msm(group, sc, bases).tobytes()
        buf = torch.zeros(n * words + 4, dtype=torch.int64, device="cuda")
        assert buf.data_ptr() % 32 == 0
        view = buf[shift // 8: shift // 8 + n * words]
        view.copy_(torch.from_numpy(bases.view(np.int64).reshape(-1)))
        assert view.data_ptr() % 32 == shift
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        got = raw_msm(b381, "g1" if group == 1 else "g2", d_sc, view, n, c=6, scalars_on_device=True, points_on_device=True)
        assert got.tobytes() == exp, (group, shift)


@pytest.mark.parametrize("levels", ["", "0", "3"])
def test_skewed_scalar_distributions(cuda, b381, oracle, monkeypatch, levels):
    """All scalars equal (commitment to a constant vector; the reference's "sum of ones" identity,
    test_msm_security.cu:410-505, at scale): every window has ONE bucket holding all n points.  The bucket is cut
    into at most 1024 tasks and summed by the warp-per-bucket finalize kernel; with affine levels it is halved
    level by level, balanced by output slot.  Checked through sum s (k0 + i d) G = s (n k0 + d n(n-1)/2) G."""
    import torch
    if levels:
        monkeypatch.setenv("B381_MSM_LEVELS", levels)
    n = 1 << 17
    rng = P.SplitMix64(0x5EED)
    k0, d, s = rng.fr(), rng.fr(), rng.fr()
    bases = oracle.gen_series(1, P.to_limbs(k0, 4), P.to_limbs(d, 4), n)
    d_bases = torch.from_numpy(bases.view(np.int64)).cuda()
    tot = (n * k0 + d * (n * (n - 1) // 2)) % P.R_MOD
    for val in (s, 1, P.R_MOD - 1):
        sc = np.tile(np.array(P.to_limbs(val, 4), dtype=np.uint64), (n, 1))
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        got = raw_msm(b381, "g1", d_sc, d_bases, n, scalars_on_device=True, points_on_device=True)
        assert got.tobytes() == P.g1_result_std_bytes(P.g1_mul(val * tot % P.R_MOD, P.G1_GEN)), val
    # half the scalars zero, the rest one of two values
    sc = np.zeros((n, 4), dtype=np.uint64)
    sc[1::2] = np.array(P.to_limbs(s, 4), dtype=np.uint64)
    sc[3::4] = np.array(P.to_limbs(7, 4), dtype=np.uint64)
    idx = np.arange(n)
    def dl(sel):
        ii = [int(i) for i in idx[sel]]
        return (len(ii) * k0 + d * sum(ii)) % P.R_MOD
    m_s = np.zeros(n, dtype=bool); m_s[1::2] = True; m_s[3::4] = False
    m_7 = np.zeros(n, dtype=bool); m_7[3::4] = True
    exp = (s * dl(m_s) + 7 * dl(m_7)) % P.R_MOD
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    got = raw_msm(b381, "g1", d_sc, d_bases, n, scalars_on_device=True, points_on_device=True)
    assert got.tobytes() == P.g1_result_std_bytes(P.g1_mul(exp, P.G1_GEN))


def test_window_sizes_agree(cuda, b381, oracle):
    """the reference never compares window sizes; we do."""
    n = 4096
    bases = oracle.gen_series(1, [3, 0, 0, 0], [5, 0, 0, 0], n)
    sc = oracle.random_fr(11, n)
    exp = oracle.msm(1, sc, bases).tobytes()
    for c in (2, 5, 9, 12, 13, 16):
        assert raw_msm(b381, "g1", sc, bases, n, c=c).tobytes() == exp, c


def test_batch_and_precompute(cuda, b381, oracle):
    """batch == individual, precompute == standard (core/msm.rs:1980-2111)."""
    n, b = 2048, 3
    bases = oracle.gen_series(1, [9, 0, 0, 0], [11, 0, 0, 0], n)
    sc = oracle.random_fr(21, n * b)
    singles = [oracle.msm(1, sc[i * n:(i + 1) * n], bases).tobytes() for i in range(b)]
    got = raw_msm(b381, "g1", sc, bases, n, batch=b)
    assert [got[i].tobytes() for i in range(b)] == singles
    # per-batch bases
    bases_b = oracle.gen_series(1, [1, 0, 0, 0], [7, 0, 0, 0], n * b)
    singles2 = [oracle.msm(1, sc[i * n:(i + 1) * n], bases_b[i * n:(i + 1) * n]).tobytes() for i in range(b)]
    got2 = raw_msm(b381, "g1", sc, bases_b, n, batch=b, shared=False)
    assert [got2[i].tobytes() for i in range(b)] == singles2
    import midnight_bls12_381_cuda_b200 as M
    ctx = M.GpuMsmContext()
    dev = ctx.upload_g1_bases(bases)
    res = ctx.msm_batch_with_device_bases([mont(sc[i * n:(i + 1) * n]) for i in range(b)], dev)
    for r, s in zip(res, singles):
        assert r[0].to_bytes(48, "little") + r[1].to_bytes(48, "little") == s[:96]
    assert ctx.msm_batch_with_device_bases_async([mont(sc[i * n:(i + 1) * n]) for i in range(b)], dev).wait() == res
    short = oracle.msm(1, sc[:777], bases[:777]).tobytes()
    for factor in (2, 4):
        pre = ctx.precompute_bases(dev, factor)
        assert pre.buffer_size() == n * factor and pre.is_precomputed()
        r = ctx.msm_with_device_bases(mont(sc[:n]), pre)
        assert r[0].to_bytes(48, "little") + r[1].to_bytes(48, "little") == singles[0][:96]
        # fewer scalars than bases over the FULL precomputed buffer (core/msm.rs:654-661), every entry point;
        # a context with another window must still use the table's own
        other = M.GpuMsmContext(window=11)
        for got in (ctx.msm_with_device_bases(mont(sc[:777]), pre), other.msm_with_device_bases(mont(sc[:777]), pre),
                    ctx.msm_with_device_bases_async(mont(sc[:777]), pre).wait(),
                    ctx.msm_batch_with_device_bases([mont(sc[:777])], pre)[0],
                    ctx.msm_batch_with_device_bases_async([mont(sc[:777])], pre).wait()[0]):
            assert got[0].to_bytes(48, "little") + got[1].to_bytes(48, "little") == short[:96]


def test_precompute_flag_combinations(cuda, b381, oracle):
    """The reference's Rust layer builds precomputed tables with are_bases_montgomery_form = true and uses them with
    false (core/msm.rs:450 vs :641-643); upstream ICICLE writes the table in the form declared for the input.  Both
    must give the right commitment, for Montgomery and for standard-form input points, on host and device buffers."""
    import torch
    lib = b381.lib()
    n, factor = 1500, 3
    bases = oracle.gen_series(1, [5, 0, 0, 0], [3, 0, 0, 0], n)          # Montgomery affine
    bases_std = np.stack([np.concatenate([oracle.unop("fq_from_mont", p[:6], 6), oracle.unop("fq_from_mont", p[6:], 6)])
                          for p in bases])
    sc = oracle.random_fr(31, n)
    exp = oracle.msm(1, sc, bases).tobytes()
    for src, declared_mont in ((bases, True), (bases_std, False)):
        cfg = lib.b381_default_msm_config()
        cfg.precompute_factor, cfg.are_points_montgomery_form = factor, declared_mont
        table = np.zeros((n * factor, 12), dtype=np.uint64)
        assert lib.b381_g1_msm_precompute_bases(b381.ptr(src), n, C.byref(cfg), b381.ptr(table)) == 0
        assert (table[::factor] == src).all()                             # multiple 0 is the point itself, same form
        d_table = torch.from_numpy(table.view(np.int64)).cuda()
        for use_flag in (declared_mont, False):
            for buf, on_dev in ((table, False), (d_table, True)):
                m = lib.b381_default_msm_config()
                m.precompute_factor, m.are_points_montgomery_form, m.are_points_on_device = factor, use_flag, on_dev
                res = np.zeros(18, dtype=np.uint64)
                assert lib.b381_g1_msm(b381.ptr(sc), b381.ptr(buf), n, C.byref(m), b381.ptr(res)) == 0
                assert res.tobytes() == exp, (declared_mont, use_flag, on_dev)


def test_forced_levels_on_sparse_buckets(cuda, b381, oracle, monkeypatch):
    """Affine levels forced onto an input with fewer points than buckets (c = 16, n = 1000): the per-level bound
    (in + nbuckets) / 2 + 1 GROWS from level to level there, so the slot-major scratch must be sized for the largest
    planned level, not for level 0 (out-of-bounds write otherwise)."""
    n = 1000
    bases = oracle.gen_series(1, [2, 0, 0, 0], [9, 0, 0, 0], n)
    sc = oracle.random_fr(77, n)
    exp = oracle.msm(1, sc, bases).tobytes()
    for levels in ("1", "3", "5"):
        monkeypatch.setenv("B381_MSM_LEVELS", levels)
        assert raw_msm(b381, "g1", sc, bases, n, c=16).tobytes() == exp, levels
        assert raw_msm(b381, "g1", sc, bases, n, c=10).tobytes() == exp, levels
    monkeypatch.delenv("B381_MSM_LEVELS")


def test_full_size_discrete_log_check(cuda, b381, oracle):
    """2^20 G1 and 2^16 G2 through the size-independent identity
    sum s_i (k_i G) = (sum s_i k_i mod r) G with P_i = (k0 + i d) G."""
    import torch
    for group, k, logn in (("g1", 1, 20), ("g2", 2, 16)):
        n = 1 << logn
        rng = P.SplitMix64(0xB12381_1016 + k)
        k0, d = rng.fr(), rng.fr()
        bases = oracle.gen_series(k, P.to_limbs(k0, 4), P.to_limbs(d, 4), n)
        sc = oracle.random_fr(0xB12381_0016 + k, n)
        # sum s_i (k0 + i d) = k0 * S0 + d * S1
        s = fr_ints(sc)
        dl = (k0 * sum(s) + d * sum(i * v for i, v in enumerate(s))) % P.R_MOD
        exp = P.g1_result_std_bytes(P.g1_mul(dl, P.G1_GEN)) if k == 1 else P.g2_result_std_bytes(P.g2_mul(dl, P.G2_GEN))
        d_bases = torch.from_numpy(bases.view(np.int64)).cuda()
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        got = raw_msm(b381, group, d_sc, d_bases, n, scalars_on_device=True, points_on_device=True)
        assert got.tobytes() == exp, group


def test_batch_is_one_folded_run(cuda, b381, oracle, monkeypatch):
    """A batch runs as ONE pipeline pass over batch * Wf bucket sets (msm_core.cuh make_msm_shape): results equal the
    individual MSMs for shared and per-MSM bases, at sizes with and without affine levels, for G2, when the batch has to
    be cut into groups (B381_MSM_BATCH_GROUP), and with host-resident scalars (chunked copy spans MSM boundaries)."""
    import torch
    lib = b381.lib()
    for (n, b, c) in [(1 << 10, 16, 0), (1 << 16, 6, 16), (3001, 5, 9)]:
        bases = oracle.gen_series(1, [3, 0, 0, 0], [5, 0, 0, 0], n)
        sc = oracle.random_fr(500 + n, n * b)
        singles = [raw_msm(b381, "g1", sc[i * n:(i + 1) * n], bases, n, c=c)[0].tobytes() for i in range(b)]
        assert singles[0] == oracle.msm(1, sc[:n], bases).tobytes()
        got = raw_msm(b381, "g1", sc, bases, n, batch=b, c=c)
        assert [got[i].tobytes() for i in range(b)] == singles, (n, b)
        monkeypatch.setenv("B381_MSM_BATCH_GROUP", "4")            # 16 -> 4 x 4, 6 -> 4 + 2, 5 -> 4 + 1
        got = raw_msm(b381, "g1", sc, bases, n, batch=b, c=c)
        monkeypatch.delenv("B381_MSM_BATCH_GROUP")
        assert [got[i].tobytes() for i in range(b)] == singles, (n, b, "groups")
    # per-MSM bases, device-resident inputs
    n, b = 1 << 12, 7
    bases_b = oracle.gen_series(1, [2, 0, 0, 0], [9, 0, 0, 0], n * b)
    sc = oracle.random_fr(77, n * b)
    singles = [oracle.msm(1, sc[i * n:(i + 1) * n], bases_b[i * n:(i + 1) * n]).tobytes() for i in range(b)]
    d_sc, d_b = torch.from_numpy(sc.view(np.int64)).cuda(), torch.from_numpy(bases_b.view(np.int64)).cuda()
    got = raw_msm(b381, "g1", d_sc, d_b, n, batch=b, shared=False, scalars_on_device=True, points_on_device=True)
    assert [got[i].tobytes() for i in range(b)] == singles
    # G2
    n2, b2 = 700, 4
    bases2 = oracle.gen_series(2, [3, 0, 0, 0], [5, 0, 0, 0], n2)
    sc2 = oracle.random_fr(78, n2 * b2)
    got2 = raw_msm(b381, "g2", sc2, bases2, n2, batch=b2)
    assert [got2[i].tobytes() for i in range(b2)] == [oracle.msm(2, sc2[i * n2:(i + 1) * n2], bases2).tobytes() for i in range(b2)]
    # empty batch element sizes
    assert raw_msm(b381, "g1", sc[:0], bases_b[:0], 0, batch=3).tobytes() == P.g1_result_std_bytes(None) * 3


@pytest.mark.parametrize("chunk_log", ["13", "15"])
def test_chunk_major_grouping(cuda, b381, oracle, monkeypatch, chunk_log):
    """msm_core.cuh CHUNK-MAJOR GROUPING (default from 2^23 points; forced here at 2^16-2^17 with small chunks): same bytes
    as the plain grouping and as the oracle; resident and host scalars (per-piece sort under the copy), batch, G2."""
    import torch
    monkeypatch.setenv("B381_MSM_LEVELS", "2")
    n = (1 << 16) + 777
    bases = oracle.gen_series(1, [3, 0, 0, 0], [5, 0, 0, 0], n)
    sc = oracle.random_fr(31, n)
    sc[5], sc[6] = 0, sc[7]
    exp = oracle.msm(1, sc, bases).tobytes()
    monkeypatch.setenv("B381_MSM_CHUNK_LOG", "31")
    assert raw_msm(b381, "g1", sc, bases, n, c=10)[0].tobytes() == exp
    monkeypatch.setenv("B381_MSM_CHUNK_LOG", chunk_log)
    assert raw_msm(b381, "g1", sc, bases, n, c=10)[0].tobytes() == exp                         # host scalars
    d_sc, d_b = torch.from_numpy(sc.view(np.int64)).cuda(), torch.from_numpy(bases.view(np.int64)).cuda()
    assert raw_msm(b381, "g1", d_sc, d_b, n, c=10, scalars_on_device=True, points_on_device=True)[0].tobytes() == exp
    # batch of 3 over shared bases; chunks straddle the MSM boundaries
    nb = 20000
    scb = oracle.random_fr(32, nb * 3)
    got = raw_msm(b381, "g1", scb, bases[:nb], nb, batch=3, c=9)
    assert [g.tobytes() for g in got] == [oracle.msm(1, scb[i * nb:(i + 1) * nb], bases[:nb]).tobytes() for i in range(3)]
    # G2
    n2 = 9000
    bases2 = oracle.gen_series(2, [3, 0, 0, 0], [5, 0, 0, 0], n2)
    sc2 = oracle.random_fr(33, n2)
    assert raw_msm(b381, "g2", sc2, bases2, n2, c=8)[0].tobytes() == oracle.msm(2, sc2, bases2).tobytes()


@pytest.mark.parametrize("chunk_log,piece_log", [("13", "14"), ("14", "14"), ("12", "15")])
def test_streamed_level0(cuda, b381, oracle, monkeypatch, chunk_log, piece_log):
    """Host scalars + chunk-major + several pieces: level 0's forward pass runs piece by piece while later pieces are
    still being copied and sorted (msm_impl.cuh `streamed`), forced here at 2^16 points with 2^14 / 2^15-scalar pieces
    of 1, 2 and 8 chunks.  Same bytes as the oracle; a zeroed piece (only trash slots), a ragged last chunk, G2."""
    monkeypatch.setenv("B381_MSM_LEVELS", "2")
    monkeypatch.setenv("B381_MSM_CHUNK_LOG", chunk_log)
    monkeypatch.setenv("B381_MSM_PIECE_LOG", piece_log)
    n = (1 << 16) + 777
    bases = oracle.gen_series(1, [3, 0, 0, 0], [5, 0, 0, 0], n)
    sc = oracle.random_fr(41, n)
    sc[5], sc[6] = 0, sc[7]
    assert raw_msm(b381, "g1", sc, bases, n, c=10)[0].tobytes() == oracle.msm(1, sc, bases).tobytes()
    sc[1 << 14:1 << 15] = 0                                            # one piece with nothing but trash slots
    assert raw_msm(b381, "g1", sc, bases, n, c=10)[0].tobytes() == oracle.msm(1, sc, bases).tobytes()
    sc[:1 << 14] = 0                                                   # ... and the first piece as well
    assert raw_msm(b381, "g1", sc, bases, n, c=9)[0].tobytes() == oracle.msm(1, sc, bases).tobytes()
    n2 = 40000
    bases2 = oracle.gen_series(2, [3, 0, 0, 0], [5, 0, 0, 0], n2)
    sc2 = oracle.random_fr(43, n2)
    assert raw_msm(b381, "g2", sc2, bases2, n2, c=8)[0].tobytes() == oracle.msm(2, sc2, bases2).tobytes()


def test_chunk_major_host_scalars_large(cuda, b381, oracle, monkeypatch):
    """2^22 host scalars in two 2^21 pieces, chunk-major with 2^20 chunks: every piece is histogrammed, scanned and
    scattered while the next one is still in flight; checked against the discrete-log identity."""
    import torch
    monkeypatch.setenv("B381_MSM_CHUNK_LOG", "20")
    lib = b381.lib()
    n = 1 << 22
    g = np.frombuffer(P.g1_affine_mont_bytes(P.G1_GEN), dtype=np.uint64).copy()
    bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    assert lib.b381_g1_point_series(b381.ptr(g), b381.ptr(g), C.c_uint64(n), b381.ptr(bases), None) == 0
    sc = oracle.random_fr(34, n)
    got = raw_msm(b381, "g1", sc, bases, n, points_on_device=True)[0].tobytes()
    kk = np.zeros((n, 4), dtype=np.uint64)
    kk[:, 0] = np.arange(1, n + 1, dtype=np.uint64)
    dl = P.from_limbs(oracle.fr_dot(sc, kk))
    assert got == P.g1_result_std_bytes(P.g1_mul(dl, P.G1_GEN))
