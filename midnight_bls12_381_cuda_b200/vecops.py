"""Fr vector ops -- mirrors core/vecops.rs:140-365 (vector_add/sub/mul, scalar_mul) over the C ABI.
The size-threshold CPU fallback of the reference (should_use_gpu_vecops, :109) is removed."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .types import TypeConverter


class VecOpsError(RuntimeError):
    """core/vecops.rs:69-103"""


def _binary(fn_name, a, b, a_is_scalar=False):
    a = np.ascontiguousarray(TypeConverter.scalar_slice_as_icicle(a))
    b = np.ascontiguousarray(TypeConverter.scalar_slice_as_icicle(b))
    if not a_is_scalar and a.shape[0] != b.shape[0]:
        raise VecOpsError(f"length mismatch: {a.shape[0]} vs {b.shape[0]}")
    out = np.empty_like(b)
    cfg = L.lib().b381_default_vecops_config()
    code = getattr(L.lib(), fn_name)(L.ptr(a), L.ptr(b), C.c_uint64(b.shape[0]), C.byref(cfg), L.ptr(out))
    if code != 0:
        raise VecOpsError(f"{fn_name}: {L.ERROR_NAMES.get(code, code)}")
    return out


def vector_add(a, b):
    return _binary("b381_vector_add", a, b)


def vector_sub(a, b):
    return _binary("b381_vector_sub", a, b)


def vector_mul(a, b):
    return _binary("b381_vector_mul", a, b)


def scalar_mul(scalar, a):
    return _binary("b381_scalar_mul_vec", scalar, a, a_is_scalar=True)


def scalar_add(scalar, a):
    return _binary("b381_scalar_add_vec", scalar, a, a_is_scalar=True)


def _unary(fn_name, a, n_out=None, extra=()):
    a = np.ascontiguousarray(TypeConverter.scalar_slice_as_icicle(a))
    n = a.shape[0]
    out = np.zeros((n if n_out is None else n_out, 4), dtype=np.uint64)
    cfg = L.lib().b381_default_vecops_config()
    code = getattr(L.lib(), fn_name)(L.ptr(a), C.c_uint64(n), *extra, C.byref(cfg), L.ptr(out))
    if code != 0:
        raise VecOpsError(f"{fn_name}: {L.ERROR_NAMES.get(code, code)}")
    return out


def bit_reverse(a):
    """core/vecops.rs:392-452 (length must be a power of two; the CPU fallback for small sizes is removed)."""
    a = np.ascontiguousarray(TypeConverter.scalar_slice_as_icicle(a))
    if a.shape[0] & (a.shape[0] - 1):
        raise VecOpsError(f"bit_reverse requires power of 2 length, got {a.shape[0]}")
    return _unary("b381_bit_reverse", a)


def bit_reverse_inplace(a):
    """core/vecops.rs:454-535: permutes the caller's numpy array."""
    a[...] = bit_reverse(a)


def vector_sum(a):
    return _unary("b381_vector_sum", a, n_out=1)[0]


def batch_inverse(a):
    return _unary("b381_vector_inv", a)


def to_montgomery(a):
    return _unary("b381_montgomery_convert", a, extra=(C.c_int(1),))


def from_montgomery(a):
    return _unary("b381_montgomery_convert", a, extra=(C.c_int(0),))
