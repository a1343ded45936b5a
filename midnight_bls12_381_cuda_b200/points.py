"""Point-level helpers over the C ABI: format conversion, ingest validation, batch scalar multiplication.

Mirrors the exported functions of bls12-381/src/curve/point_ops.cu:752-1268 (same names on the C side) plus the two
subgroup checks the reference leaves as TODO (include/point.cuh:419-448).  Points are numpy uint64 arrays in the wire
layouts of types.py: G1 affine [n, 12], G1 Jacobian [n, 18], G2 affine [n, 24], G2 Jacobian [n, 36] (Montgomery)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class PointError(RuntimeError):
    pass


def _call(name: str, inp: np.ndarray, n: int, out: np.ndarray, *more) -> np.ndarray:
    cfg = L.lib().b381_default_vecops_config()
    args = [L.ptr(np.ascontiguousarray(inp))] + [L.ptr(np.ascontiguousarray(m)) for m in more]
    code = getattr(L.lib(), name)(*args, n, C.byref(cfg), L.ptr(out))
    if code != 0:
        raise PointError(f"{name}: {L.ERROR_NAMES.get(code, code)}")
    return out


def _rows(a, words: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, words)
    if a.shape[0] == 0:
        raise PointError("empty input")
    return a


def g1_affine_to_projective(points) -> np.ndarray:
    p = _rows(points, 12)
    return _call("bls12_381_g1_affine_to_projective", p, p.shape[0], np.empty((p.shape[0], 18), dtype=np.uint64))


def g1_projective_to_affine(points) -> np.ndarray:
    """one shared inversion per 16 points (Montgomery's trick); the reference inverts every Z on its own"""
    p = _rows(points, 18)
    return _call("bls12_381_g1_projective_to_affine", p, p.shape[0], np.empty((p.shape[0], 12), dtype=np.uint64))


def g2_affine_to_projective(points) -> np.ndarray:
    p = _rows(points, 24)
    return _call("bls12_381_g2_affine_to_projective", p, p.shape[0], np.empty((p.shape[0], 36), dtype=np.uint64))


def g2_projective_to_affine(points) -> np.ndarray:
    p = _rows(points, 36)
    return _call("bls12_381_g2_projective_to_affine", p, p.shape[0], np.empty((p.shape[0], 24), dtype=np.uint64))


def g1_is_on_curve(points) -> np.ndarray:
    p = _rows(points, 12)
    return _call("b381_g1_is_on_curve", p, p.shape[0], np.empty(p.shape[0], dtype=np.uint8)).astype(bool)


def g2_is_on_curve(points) -> np.ndarray:
    p = _rows(points, 24)
    return _call("b381_g2_is_on_curve", p, p.shape[0], np.empty(p.shape[0], dtype=np.uint8)).astype(bool)


def g1_is_in_subgroup(points) -> np.ndarray:
    """phi(P) == [z^2 - 1] P for points already known to be on the curve; infinity counts as a member"""
    p = _rows(points, 12)
    return _call("b381_g1_is_in_subgroup", p, p.shape[0], np.empty(p.shape[0], dtype=np.uint8)).astype(bool)


def g2_is_in_subgroup(points) -> np.ndarray:
    """psi(P) == [z] P"""
    p = _rows(points, 24)
    return _call("b381_g2_is_in_subgroup", p, p.shape[0], np.empty(p.shape[0], dtype=np.uint8)).astype(bool)


def validate_g1_bases(points) -> None:
    """ingest check for externally supplied bases: on the curve AND in the order-r subgroup"""
    p = _rows(points, 12)
    ok = g1_is_on_curve(p)
    if not ok.all():
        raise PointError(f"{int((~ok).sum())} point(s) not on the curve, first at index {int(np.argmin(ok))}")
    ok = g1_is_in_subgroup(p)
    if not ok.all():
        raise PointError(f"{int((~ok).sum())} point(s) outside the prime-order subgroup, first at index {int(np.argmin(ok))}")


def g1_scalar_mul(bases, scalars, glv: bool = True) -> np.ndarray:
    """out[i] = scalars[i] * bases[i]; scalars are CANONICAL integers ([n, 4] uint64), as in the reference's kernels
    (point_ops.cu:364, :491).  Returns normalised Jacobian points (x, y, 1) / (0, R, 0)."""
    p, s = _rows(bases, 12), _rows(scalars, 4)
    if p.shape[0] != s.shape[0]:
        raise PointError(f"length mismatch: {p.shape[0]} bases, {s.shape[0]} scalars")
    name = "bls12_381_g1_scalar_mul_glv" if glv else "bls12_381_g1_scalar_mul"
    return _call(name, p, p.shape[0], np.empty((p.shape[0], 18), dtype=np.uint64), s)
