"""ctypes wrapper of oracle/oracle.c (TEST INFRASTRUCTURE ONLY -- see the header of oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _u64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a if shape is None else a.reshape(shape)


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def num_threads() -> int:
    return lib().orc_num_threads()


def constants():
    out = np.zeros(36, dtype=np.uint64)
    lib().orc_constants(_p(out))
    return out


def fr_binop(name: str, a, b):
    a, b = _u64(a), _u64(b)
    r = np.zeros(4, dtype=np.uint64)
    getattr(lib(), f"orc_fr_{name}")(_p(a), _p(b), _p(r))
    return r


def fq_binop(name: str, a, b):
    a, b = _u64(a), _u64(b)
    r = np.zeros(6, dtype=np.uint64)
    getattr(lib(), f"orc_fq_{name}")(_p(a), _p(b), _p(r))
    return r


def unop(name: str, a, n):
    a = _u64(a)
    r = np.zeros(n, dtype=np.uint64)
    getattr(lib(), f"orc_{name}")(_p(a), _p(r))
    return r


def vecop(op: int, a, b, a_scalar=False):
    a, b = _u64(a), _u64(b)
    n = b.size // 4
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_vecop(op, int(a_scalar), _p(a), _p(b), C.c_size_t(n), _p(out))
    return out


def generator(k: int):
    out = np.zeros(12 * k, dtype=np.uint64)
    lib().orc_generator(k, _p(out))
    return out


def scalar_mul(k: int, s, p):
    s, p = _u64(s), _u64(p)
    out = np.zeros(18 * k, dtype=np.uint64)
    lib().orc_scalar_mul(k, _p(s), _p(p), _p(out))
    return out


def gen_series(k: int, k0, d, n: int):
    """points (k0 + i d) G, Montgomery affine, shape (n, 12k)."""
    k0, d = _u64(k0), _u64(d)
    out = np.zeros((n, 12 * k), dtype=np.uint64)
    lib().orc_gen_series(k, _p(k0), _p(d), C.c_size_t(n), _p(out))
    return out


def on_curve(k: int, p) -> bool:
    p = _u64(p)
    return bool(lib().orc_on_curve(k, _p(p)))


def msm(k: int, scalars, bases, scalars_mont=False, c=0):
    scalars, bases = _u64(scalars), _u64(bases)
    n = scalars.size // 4
    assert bases.size == n * 12 * k
    out = np.zeros(18 * k, dtype=np.uint64)
    rc = lib().orc_msm(k, _p(scalars), int(scalars_mont), _p(bases), C.c_size_t(n), c, _p(out))
    assert rc == 0
    return out


def fr_dot(s, kk, s_mont=False):
    s, kk = _u64(s), _u64(kk)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_fr_dot(_p(s), _p(kk), C.c_size_t(s.size // 4), int(s_mont), _p(out))
    return out


def omega(log_n: int):
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_omega(log_n, _p(out))
    return out


def ntt(a, inverse=False):
    """natural in / natural out on a COPY; a: (n,4) Montgomery."""
    a = np.array(_u64(a), copy=True).reshape(-1, 4)
    log_n = (a.shape[0]).bit_length() - 1
    assert 1 << log_n == a.shape[0]
    assert lib().orc_ntt(_p(a), log_n, int(inverse)) == 0
    return a


def coset_ntt(a, g, inverse=False):
    a = np.array(_u64(a), copy=True).reshape(-1, 4)
    g = _u64(g)
    log_n = (a.shape[0]).bit_length() - 1
    assert lib().orc_coset_ntt(_p(a), log_n, int(inverse), _p(g)) == 0
    return a


def poly_eval(a, z):
    """sum_j a[j] z^j (a: (n,4) Montgomery, z: Montgomery limbs) -> Montgomery limbs; Horner, no NTT code involved."""
    a, z = _u64(a), _u64(z)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_fr_poly_eval(_p(a), C.c_size_t(a.size // 4), _p(z), _p(out))
    return out


def bit_reverse(a):
    a = np.array(_u64(a), copy=True).reshape(-1, 4)
    lib().orc_bit_reverse(_p(a), (a.shape[0]).bit_length() - 1)
    return a


def random_fr(seed: int, n: int):
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_random_fr(C.c_uint64(seed), C.c_size_t(n), _p(out))
    return out
