"""Same inputs through the REFERENCE's own CUDA kernels (oracle/_ref/libref_field.so, built by
`make -C oracle ref` from /root/reference/bls12-381/src/field/{ntt_kernels,vec_ops}.cu, unmodified)
and through this backend: identical bytes.

The reference's domain init assumes the root it is given has order 2^24 and is Montgomery
(ntt_kernels.cu:1614-1644); passing exactly that makes its tables right for every size <= 2^24, which
is the only setting in which its NTT output is meaningful.  Its registered path ignores ordering and
coset, so the comparison is kNN forward/inverse (+ batch) and the three vector ops."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyref as P
from vectors import fr_array

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_field.so")


@pytest.fixture(scope="module")
def ref(cuda, b381):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_field.so not built (needs /root/reference at build time)")
    lib = C.CDLL(REF_SO)
    root = fr_array([P.fr_to_mont(P.fr_omega(24))])
    cfg = b381.NTTInitDomainConfig()
    assert lib.bls12_381_ntt_init_domain_cuda(b381.ptr(root), C.byref(cfg)) == 0
    yield lib
    lib.bls12_381_ntt_release_domain_cuda()


@pytest.mark.parametrize("logn", [1, 3, 5, 8, 9, 12, 16, 20])
def test_ntt_same_bytes_as_reference_kernels(ref, b381, oracle, logn):
    import midnight_bls12_381_cuda_b200 as M
    ctx = M.GpuNttContext(24)
    n = 1 << logn
    a = oracle.random_fr(1000 + logn, n)
    cfg = b381.lib().b381_default_ntt_config()
    for direction in (0, 1):
        theirs = np.empty_like(a)
        assert ref.bls12_381_ntt_cuda(b381.ptr(a), n, direction, C.byref(cfg), b381.ptr(theirs)) == 0
        ours = ctx.forward_ntt(a) if direction == 0 else ctx.inverse_ntt(a)
        assert (ours == theirs).all(), (logn, direction)


def test_batched_ntt_same_bytes(ref, b381, oracle):
    import midnight_bls12_381_cuda_b200 as M
    ctx = M.GpuNttContext(24)
    for logn, batch in ((4, 16), (8, 5), (10, 3)):
        n = 1 << logn
        a = oracle.random_fr(77 + logn, n * batch)
        cfg = b381.lib().b381_default_ntt_config()
        cfg.batch_size = batch
        theirs = np.empty_like(a)
        assert ref.bls12_381_ntt_cuda(b381.ptr(a), n, 0, C.byref(cfg), b381.ptr(theirs)) == 0
        assert (ctx.forward_ntt_batch(a, n) == theirs).all(), (logn, batch)


def test_vecops_same_bytes(ref, b381, oracle):
    import midnight_bls12_381_cuda_b200 as M
    n = 65536
    a, b = oracle.random_fr(5, n), oracle.random_fr(6, n)
    cfg = b381.lib().b381_default_vecops_config()
    for name, ours in (("add", M.vecops.vector_add), ("sub", M.vecops.vector_sub), ("mul", M.vecops.vector_mul)):
        theirs = np.empty_like(a)
        fn = getattr(ref, f"bls12_381_vector_{name}")
        assert fn(b381.ptr(a), b381.ptr(b), C.c_uint64(n), C.byref(cfg), b381.ptr(theirs)) == 0
        assert (ours(a, b) == theirs).all(), name


REF_MSM_SO = os.path.join(os.path.dirname(REF_SO), "libref_msm.so")


@pytest.mark.parametrize("logn", [10, 14, 16])
def test_g1_msm_same_point_as_reference_kernels(cuda, b381, oracle, logn):
    """The reference's OWN G1 MSM (msm::msm_cuda, src/curve/msm_kernels.cu:603-903, compiled by `make -C oracle ref_msm`)
    and this backend on the same integer-form scalars and Montgomery bases, both through the flat signature
    bls12_381_g1_msm_cuda (icicle_curve_api.cu:679-692).  Both answer in Jacobian Montgomery coordinates with their own Z,
    so the comparison is on the affine point; ours must also equal the oracle."""
    if not os.path.exists(REF_MSM_SO):
        pytest.skip("oracle/_ref/libref_msm.so not built (needs /root/reference at build time)")
    ref = C.CDLL(REF_MSM_SO)
    lib = b381.lib()
    n = 1 << logn
    bases = oracle.gen_series(1, [logn, 0, 0, 0], [3, 0, 0, 0], n)
    sc = oracle.random_fr(0xB12381_5000 + logn, n)            # integer form
    cfg = lib.b381_default_msm_config()
    theirs, ours = np.zeros(18, dtype=np.uint64), np.zeros(18, dtype=np.uint64)
    assert ref.bls12_381_g1_msm_cuda(b381.ptr(sc), b381.ptr(bases), n, C.byref(cfg), b381.ptr(theirs)) == 0
    assert lib.bls12_381_g1_msm_cuda(b381.ptr(sc), b381.ptr(bases), n, C.byref(cfg), b381.ptr(ours)) == 0

    def affine(j):
        x, y, z = (P.fq_from_mont(P.from_limbs(j[6 * i:6 * i + 6])) for i in range(3))
        if z == 0:
            return None
        zi = pow(z, -1, P.P_MOD)
        return (x * zi * zi % P.P_MOD, y * zi * zi * zi % P.P_MOD)
    assert affine(theirs) == affine(ours)
    exp = oracle.msm(1, sc, bases)                            # ICICLE standard form (x, y, 1)
    assert affine(ours) == (P.from_limbs(exp[0:6]), P.from_limbs(exp[6:12]))
