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
