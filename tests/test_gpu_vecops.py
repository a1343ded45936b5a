"""Fr vecops parity on the GPU (test_vec_ops.cu: identities + reference cross-check, 64K and 1000)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyref as P
from vectors import fr_array, fr_ints, load_golden

pytestmark = pytest.mark.gpu
GOLD = load_golden()


def test_golden(cuda):
    import midnight_bls12_381_cuda_b200 as M
    v = GOLD["vecops"][0]
    a = fr_array([int(h, 16) for h in v["a"]])
    b = fr_array([int(h, 16) for h in v["b"]])
    assert [hex(x) for x in fr_ints(M.vecops.vector_add(a, b))] == v["add"]
    assert [hex(x) for x in fr_ints(M.vecops.vector_sub(a, b))] == v["sub"]
    assert [hex(x) for x in fr_ints(M.vecops.vector_mul(a, b))] == v["mul"]
    assert [hex(x) for x in fr_ints(M.vecops.scalar_mul(a[0], b))] == v["scalar_mul"]
    assert [hex(x) for x in fr_ints(M.vecops.scalar_add(a[0], b))] == v["scalar_add"]


@pytest.mark.parametrize("n", [1, 1000, 65536, (1 << 20) + 3])
def test_vs_oracle(cuda, oracle, n):
    import midnight_bls12_381_cuda_b200 as M
    a, b = oracle.random_fr(n, n), oracle.random_fr(n + 1, n)
    assert (M.vecops.vector_add(a, b) == oracle.vecop(0, a, b)).all()
    assert (M.vecops.vector_sub(a, b) == oracle.vecop(1, a, b)).all()
    assert (M.vecops.vector_mul(a, b) == oracle.vecop(2, a, b)).all()
    assert (M.vecops.scalar_mul(a[0], b) == oracle.vecop(2, a[0], b, a_scalar=True)).all()
    assert (M.vecops.scalar_add(a[0], b) == oracle.vecop(0, a[0], b, a_scalar=True)).all()


def test_identities_device_buffers_and_errors(cuda, b381, oracle):
    lib = b381.lib()
    n = 4096
    a = oracle.random_fr(9, n)
    one = np.tile(np.array(P.to_limbs(P.FR_R, 4), dtype=np.uint64), (n, 1))
    zero = np.zeros((n, 4), dtype=np.uint64)
    import midnight_bls12_381_cuda_b200 as M
    assert (M.vecops.vector_mul(a, one) == a).all() and (M.vecops.vector_add(a, zero) == a).all()
    assert not M.vecops.vector_sub(a, a).any()
    da, db = cuda.from_numpy(a.view(np.int64)).cuda(), cuda.from_numpy(one.view(np.int64)).cuda()
    out = cuda.empty_like(da)
    cfg = lib.b381_default_vecops_config()
    cfg.is_a_on_device = cfg.is_b_on_device = cfg.is_result_on_device = True
    assert lib.bls12_381_vector_mul(b381.ptr(da), b381.ptr(db), C.c_uint64(n), C.byref(cfg), b381.ptr(out)) == 0
    assert cuda.equal(out, da)
    assert lib.b381_vector_add(None, None, C.c_uint64(4), C.byref(cfg), None) == 3
    assert lib.b381_vector_add(None, None, C.c_uint64(0), C.byref(cfg), None) == 0
    with pytest.raises(M.vecops.VecOpsError):
        M.vecops.vector_add(a, a[:5])


@pytest.mark.parametrize("n", [1, 2, 1000, 1 << 16, (1 << 18) + 5])
def test_sum_inverse_and_montgomery_conversion(cuda, n):
    """vector sum, batched inversion (0 -> 0, reference field_inv convention) and the
    standard <-> Montgomery conversion, against Python big integers on the same seeded inputs."""
    import midnight_bls12_381_cuda_b200 as M
    rng = P.SplitMix64(1234 + n)
    vals = [rng.fr() for _ in range(min(n, 3000))]
    vals = (vals * (n // len(vals) + 1))[:n]
    if n > 4:
        vals[1], vals[3], vals[-1] = 0, P.R_MOD - 1, 1
    a = fr_array([P.fr_to_mont(v) for v in vals])
    assert fr_ints(M.vecops.vector_sum(a).reshape(1, 4)) == [P.fr_to_mont(sum(vals) % P.R_MOD)]
    inv_small = {v: (pow(v, -1, P.R_MOD) if v else 0) for v in set(vals)}
    assert fr_ints(M.vecops.batch_inverse(a)) == [P.fr_to_mont(inv_small[v]) for v in vals]
    assert fr_ints(M.vecops.from_montgomery(a)) == vals
    assert (M.vecops.to_montgomery(fr_array(vals)) == a).all()


@pytest.mark.parametrize("log_n", [0, 1, 3, 10, 17])
def test_bit_reverse(cuda, b381, oracle, log_n):
    """core/vecops.rs:392-535: out[bitrev(i)] = in[i]; [a0..a7] -> [a0,a4,a2,a6,a1,a5,a3,a7]; in place on the device."""
    import midnight_bls12_381_cuda_b200 as M
    n = 1 << log_n
    a = oracle.random_fr(77 + log_n, n)
    rev = np.array([int(format(i, f"0{log_n}b")[::-1], 2) if log_n else 0 for i in range(n)])
    exp = np.empty_like(a)
    exp[rev] = a
    assert (M.vecops.bit_reverse(a) == exp).all()
    if log_n == 3:
        assert [int(x[0]) for x in M.vecops.bit_reverse(fr_array(list(range(8))))] == [0, 4, 2, 6, 1, 5, 3, 7]
    lib = b381.lib()
    d = cuda.from_numpy(a.view(np.int64)).cuda()
    cfg = lib.b381_default_vecops_config()
    cfg.is_a_on_device = cfg.is_result_on_device = True
    assert lib.b381_bit_reverse(b381.ptr(d), C.c_uint64(n), C.byref(cfg), b381.ptr(d)) == 0
    assert (d.cpu().numpy().view(np.uint64) == exp).all()
    with pytest.raises(M.vecops.VecOpsError):
        M.vecops.bit_reverse(oracle.random_fr(1, 6))
    assert lib.b381_bit_reverse(b381.ptr(d), C.c_uint64(6), C.byref(cfg), b381.ptr(d)) == 11     # INVALID_ARGUMENT


def test_reference_named_flat_vec_ops(cuda, b381, oracle):
    """vec_add_cuda & co. (bls12-381/src/field/vec_ops.cu:393-520): device vectors, output first, host scalar."""
    lib = b381.lib()
    n = 5000
    a, b = oracle.random_fr(5, n), oracle.random_fr(6, n)
    da, db = cuda.from_numpy(a.view(np.int64)).cuda(), cuda.from_numpy(b.view(np.int64)).cuda()
    out = cuda.empty_like(da)
    cfg = lib.b381_default_vecops_config()
    for name, op in (("vec_add_cuda", 0), ("vec_sub_cuda", 1), ("vec_mul_cuda", 2)):
        assert getattr(lib, name)(b381.ptr(out), b381.ptr(da), b381.ptr(db), n, C.byref(cfg)) == 0
        assert (out.cpu().numpy().view(np.uint64) == oracle.vecop(op, a, b)).all(), name
    s = np.ascontiguousarray(a[7])
    assert lib.scalar_mul_vec_cuda(b381.ptr(out), b381.ptr(s), b381.ptr(db), n, C.byref(cfg)) == 0
    assert (out.cpu().numpy().view(np.uint64) == oracle.vecop(2, s, b, a_scalar=True)).all()
    assert lib.scalar_add_vec_cuda(b381.ptr(out), b381.ptr(s), b381.ptr(db), n, C.byref(cfg)) == 0
    assert (out.cpu().numpy().view(np.uint64) == oracle.vecop(0, s, b, a_scalar=True)).all()
    res = np.zeros(4, dtype=np.uint64)
    assert lib.vec_sum_cuda(b381.ptr(res), b381.ptr(da), n, C.byref(cfg)) == 0
    assert fr_ints(res.reshape(1, 4)) == [sum(fr_ints(a)) % P.R_MOD]


def test_scalar_vec_batch(cuda, b381, oracle):
    """b381_scalar_{mul,add}_vec_batch: upstream ICICLE v4 batch semantics (batch_size scalars, batch_size vectors stored
    as rows or interleaved as columns)"""
    lib = b381.lib()
    n, batch = 777, 4
    b = oracle.random_fr(61, n * batch)
    s = oracle.random_fr(62, batch)
    cfg = lib.b381_default_vecops_config()
    for name, op in (("b381_scalar_mul_vec_batch", 2), ("b381_scalar_add_vec_batch", 0)):
        for columns in (False, True):
            out = np.empty_like(b)
            assert getattr(lib, name)(b381.ptr(s), b381.ptr(b), C.c_uint64(n), batch, columns, C.byref(cfg), b381.ptr(out)) == 0
            for k in range(batch):
                got = out[k::batch] if columns else out[k * n:(k + 1) * n]
                src = np.ascontiguousarray(b[k::batch]) if columns else b[k * n:(k + 1) * n]
                assert (got == oracle.vecop(op, s[k:k + 1], src, a_scalar=True)).all(), (name, columns, k)
        assert getattr(lib, name)(b381.ptr(s), b381.ptr(b), C.c_uint64(n), 0, False, C.byref(cfg), b381.ptr(out)) == 11
