"""Fr NTT parity on the GPU through the C ABI: golden vectors (every ordering x direction x coset),
oracle.c on seeded inputs up to 2^20, the reference's property tests (test_ntt_security.cu:993-1013,
core/ntt.rs:2005-2305) and size-independent checks at 2^24."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import pyref as P
from vectors import ORDERINGS, fr_array, fr_ints, load_golden, ntt_case_input

pytestmark = pytest.mark.gpu
GOLD = load_golden()


@pytest.fixture(scope="module")
def ctx(cuda):
    import midnight_bls12_381_cuda_b200 as M
    return M.GpuNttContext(24)


def mont(vals):
    return fr_array([P.fr_to_mont(v) for v in vals])


@pytest.mark.parametrize("case", GOLD["ntt"], ids=[c["name"] for c in GOLD["ntt"]])
def test_golden(ctx, case):
    import midnight_bls12_381_cuda_b200 as M
    a = mont(ntt_case_input(case))
    c = M.GpuNttContext(24, ordering=ORDERINGS[case["ordering"]])
    g = mont([case["coset"]])[0] if case["coset"] else None
    if g is None:
        out = c.inverse_ntt(a) if case["inverse"] else c.forward_ntt(a)
    else:
        out = c.inverse_coset_ntt(a, g) if case["inverse"] else c.forward_coset_ntt(a, g)
    assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"]
    assert [hex(v) for v in fr_ints(out[: len(case["head"])])] == case["head"]
    b = a.copy()                      # in place gives the same bytes
    if g is None:
        (c.inverse_ntt_inplace if case["inverse"] else c.forward_ntt_inplace)(b)
    else:
        (c.inverse_coset_ntt_inplace if case["inverse"] else c.forward_coset_ntt_inplace)(b, g)
    assert (b == out).all()


@pytest.mark.parametrize("logn", [1, 2, 5, 8, 11, 12, 15, 20])
def test_vs_oracle(ctx, oracle, logn):
    """config 2 of BASELINE.json: NTT + iNTT up to 2^20, natural and coset, vs the CPU path."""
    a = oracle.random_fr(0xB12381_2020 + logn, 1 << logn)
    y = ctx.forward_ntt(a)
    assert (y == oracle.ntt(a)).all()
    assert (ctx.inverse_ntt(y) == a).all()
    g = mont([7])[0]
    yc = ctx.forward_coset_ntt(a, g)
    assert (yc == oracle.coset_ntt(a, g)).all()
    assert (ctx.inverse_coset_ntt(yc, g) == a).all()


def test_orderings_on_device(ctx, oracle, cuda):
    import midnight_bls12_381_cuda_b200 as M
    n = 1 << 13
    a = oracle.random_fr(5, n)
    nat = oracle.ntt(a)
    d = cuda.from_numpy(a.view(np.int64)).cuda()
    for name, code in ORDERINGS.items():
        src = oracle.bit_reverse(a) if name[0] == "R" else a
        exp = oracle.bit_reverse(nat) if name[1] == "R" else nat
        t = cuda.from_numpy(src.view(np.int64)).cuda()
        ctx.ntt_on_device(t.data_ptr(), M.ntt.FORWARD, size=n, ordering=code)
        assert (t.cpu().numpy().view(np.uint64).reshape(-1, 4) == exp).all(), name
    # NR then RN-inverse is the no-reorder round trip (core/ntt.rs:262-280)
    ctx.ntt_on_device(d.data_ptr(), M.ntt.FORWARD, size=n, ordering=M.ntt.kNR)
    ctx.ntt_on_device(d.data_ptr(), M.ntt.INVERSE, size=n, ordering=M.ntt.kRN)
    assert (d.cpu().numpy().view(np.uint64).reshape(-1, 4) == a).all()


def test_batch_and_columns(ctx, oracle, b381):
    """batch == individual (core/ntt.rs:2160-2200), incl. warp-/smem-sized transforms and columns_batch."""
    for logn, batch in ((3, 64), (5, 33), (8, 7), (12, 3)):
        n = 1 << logn
        a = oracle.random_fr(logn * 100 + batch, n * batch)
        exp = np.concatenate([oracle.ntt(a[i * n:(i + 1) * n]) for i in range(batch)])
        assert (ctx.forward_ntt_batch(a, n) == exp).all(), (logn, batch)
        assert (ctx.inverse_ntt_batch(exp, n) == a).all()
        # columns_batch: element j of transform b at j*batch + b
        cols = a.reshape(batch, n, 4).transpose(1, 0, 2).reshape(-1, 4).copy()
        out = np.empty_like(cols)
        cfg = b381.lib().b381_default_ntt_config()
        cfg.batch_size, cfg.columns_batch = batch, True
        assert b381.lib().b381_ntt(b381.ptr(cols), n, 0, C.byref(cfg), b381.ptr(out)) == 0
        assert (out.reshape(n, batch, 4).transpose(1, 0, 2).reshape(-1, 4) == exp).all()


def test_properties(ctx, oracle):
    n = 1 << 10
    a, b = oracle.random_fr(1, n), oracle.random_fr(2, n)
    A, B = ctx.forward_ntt(a), ctx.forward_ntt(b)
    assert not ctx.forward_ntt(np.zeros((n, 4), dtype=np.uint64)).any()                     # zeros -> zeros
    import midnight_bls12_381_cuda_b200 as M
    assert (ctx.forward_ntt(M.vecops.vector_add(a, b)) == M.vecops.vector_add(A, B)).all()  # linearity
    delta = np.zeros((n, 4), dtype=np.uint64)
    delta[0] = P.to_limbs(P.FR_R, 4)
    assert (ctx.forward_ntt(delta) == np.tile(delta[0], (n, 1))).all()                      # delta -> ones
    conv = ctx.inverse_ntt(M.vecops.vector_mul(A, B))                                        # convolution theorem
    ai, bi = [P.fr_from_mont(v) for v in fr_ints(a)], [P.fr_from_mont(v) for v in fr_ints(b)]
    for i in (0, 1, 517, n - 1):
        assert P.fr_from_mont(fr_ints(conv[i])[0]) == sum(ai[j] * bi[(i - j) % n] for j in range(n)) % P.R_MOD
    assert (ctx.forward_ntt(a) == A).all()                                                   # determinism
    assert (ctx.forward_ntt_async(a).wait() == A).all()


def test_errors_and_domain(ctx, b381, oracle):
    lib = b381.lib()
    cfg = lib.b381_default_ntt_config()
    a = oracle.random_fr(3, 12)
    assert lib.b381_ntt(b381.ptr(a), 12, 0, C.byref(cfg), b381.ptr(a)) == 11       # not a power of two
    assert lib.b381_ntt(None, 8, 0, C.byref(cfg), None) == 3
    rou = np.zeros(4, dtype=np.uint64)
    for k in (0, 1, 2, 10, 24):
        assert lib.b381_ntt_get_rou_from_domain(k, b381.ptr(rou)) == 0
        assert P.from_limbs(rou) == P.fr_to_mont(P.fr_omega(k))
    assert lib.b381_ntt_get_rou_from_domain(25, b381.ptr(rou)) == 11
    # Montgomery-form root (what the reference's CUDA tests pass) is recognised as such
    import midnight_bls12_381_cuda_b200.ntt as N
    lib.b381_ntt_release_domain()
    N._domain_log = 0
    assert lib.b381_ntt(b381.ptr(a[:8]), 8, 0, C.byref(cfg), b381.ptr(np.empty((8, 4), dtype=np.uint64))) == 11  # no domain
    root_m = fr_array([P.fr_to_mont(P.fr_omega(16))])
    assert lib.bls12_381_ntt_init_domain_cuda(b381.ptr(root_m), C.byref(b381.NTTInitDomainConfig())) == 0
    out = np.empty((8, 4), dtype=np.uint64)
    assert lib.bls12_381_ntt_cuda(b381.ptr(a[:8].copy()), 8, 0, C.byref(cfg), b381.ptr(out)) == 0
    assert (out == oracle.ntt(a[:8])).all()
    assert lib.b381_ntt_get_rou_from_domain(17, b381.ptr(rou)) == 11
    bad = fr_array([12345])
    lib.b381_ntt_release_domain()
    assert lib.b381_ntt_init_domain(b381.ptr(bad), C.byref(b381.NTTInitDomainConfig())) == 11   # not a 2-power root
    N.GpuNttContext(24)       # restore for later tests


@pytest.mark.parametrize("logn", [21, 22, 24])
def test_natural_order_equals_bit_reversed_order(ctx, oracle, cuda, logn):
    """Large transforms in both output orders: kNN scatters the last pass's stores to bit-reversed addresses, kNR
    stores in place.  Same transform, so y_NN[i] == y_NR[bitrev(i)] element for element, forward and inverse."""
    import midnight_bls12_381_cuda_b200 as M
    n = 1 << logn
    a = oracle.random_fr(0xB12381_3000 + logn, n)
    idx = cuda.arange(n, dtype=cuda.int64, device="cuda")
    rev = cuda.zeros_like(idx)
    for b in range(logn):
        rev |= ((idx >> b) & 1) << (logn - 1 - b)
    del idx
    for direction in (M.ntt.FORWARD, M.ntt.INVERSE):
        nn = cuda.from_numpy(a.view(np.int64)).cuda()
        nr = nn.clone()
        ctx.ntt_on_device(nn.data_ptr(), direction, size=n, ordering=M.ntt.kNN)
        ctx.ntt_on_device(nr.data_ptr(), direction, size=n, ordering=M.ntt.kNR)
        assert cuda.equal(nn, nr[rev]), (logn, direction)
        del nn, nr


@pytest.mark.parametrize("logn", [21, 22, 24])
def test_large_vs_oracle(ctx, oracle, cuda, logn):
    """BASELINE.json's headline NTT size and its neighbours (the multi-pass plans): every output byte of the forward
    kNN transform, the inverse and the coset-7 pair against oracle.ntt / oracle.coset_ntt (core/ntt.rs:1488-1603
    contract: forward = best_fft, inverse(forward(x)) = x, coset = scale by g^i first)."""
    import midnight_bls12_381_cuda_b200 as M
    n = 1 << logn
    a = oracle.random_fr(0xB12381_2100 + logn, n)
    d = cuda.from_numpy(a.view(np.int64)).cuda()

    def host(t):
        return t.cpu().numpy().view(np.uint64).reshape(-1, 4)
    ctx.ntt_on_device(d.data_ptr(), M.ntt.FORWARD, size=n)
    y = host(d)
    exp = oracle.ntt(a)
    assert (y == exp).all(), f"forward 2^{logn}: first mismatch at {int(np.argmax((y != exp).any(axis=1)))}"
    # inverse of an independent vector (not just the round trip): oracle.intt(b) for random b
    b = oracle.random_fr(0xB12381_2200 + logn, n)
    d.copy_(cuda.from_numpy(b.view(np.int64)))
    ctx.ntt_on_device(d.data_ptr(), M.ntt.INVERSE, size=n)
    assert (host(d) == oracle.ntt(b, inverse=True)).all(), f"inverse 2^{logn}"
    g = mont([7])[0]
    yc = ctx.forward_coset_ntt(a, g)
    assert (yc == oracle.coset_ntt(a, g)).all(), f"coset forward 2^{logn}"
    assert (ctx.inverse_coset_ntt(yc, g) == a).all(), f"coset inverse 2^{logn}"


def test_full_size_2_24(ctx, oracle, cuda):
    """2^24: inverse(forward(x)) == x, y[0] == sum, and 32 random outputs re-evaluated as polynomial values
    y[i] = sum_j a[j] w^(ij) with big integers (spot checks that do not share code with any NTT)."""
    import random

    import midnight_bls12_381_cuda_b200 as M
    n = 1 << 24
    a = oracle.random_fr(0xB12381_2024, n)
    d = cuda.from_numpy(a.view(np.int64)).cuda()
    ctx.ntt_on_device(d.data_ptr(), M.ntt.FORWARD, size=n)
    y = d.cpu().numpy().view(np.uint64).reshape(-1, 4)
    # i = 0 is the plain sum; check it exactly with limb-wise sums of the Montgomery words
    lo = (a & np.uint64(0xFFFFFFFF)).sum(axis=0, dtype=np.uint64)
    hi = (a >> np.uint64(32)).sum(axis=0, dtype=np.uint64)
    tot = sum((int(lo[l]) + (int(hi[l]) << 32)) << (64 * l) for l in range(4)) % P.R_MOD
    assert P.from_limbs(y[0]) == tot
    # y[i] = A(w^i): Horner evaluation of the input as a polynomial (oracle.poly_eval), 32 positions
    w = P.fr_omega(24)
    rng = random.Random(0xB12381)
    for i in [1, n // 2 + 12345, n - 1] + [rng.randrange(n) for _ in range(29)]:
        z = fr_array([P.fr_to_mont(pow(w, i, P.R_MOD))])[0]
        assert (y[i] == oracle.poly_eval(a, z)).all(), i
    ctx.ntt_on_device(d.data_ptr(), M.ntt.INVERSE, size=n)
    assert cuda.equal(d, cuda.from_numpy(a.view(np.int64)).cuda())
