"""TypeConverter -- the host layout contract of core/types.rs:89-108,126-450.

midnight_curves::Fq (the SCALAR field, called `Scalar` in the reference) == [u64; 4] little-endian
Montgomery == ICICLE ScalarField bytes; Fp == [u64; 6]; G1Affine == x||y (96 B); Fp2 == c0||c1;
G2Affine 192 B.  Views are zero-copy (numpy reshape/view), exactly like the reference's transmutes.
Results come back as ICICLE projective points in STANDARD form: (x, y, 1) or (0, 1, 0).
"""
from __future__ import annotations

import numpy as np

FR_BYTES, FQ_BYTES = 32, 48
G1_AFFINE_BYTES, G1_PROJECTIVE_BYTES = 96, 144
G2_AFFINE_BYTES, G2_PROJECTIVE_BYTES = 192, 288


class TypeConverter:
    @staticmethod
    def scalar_slice_as_icicle(scalars) -> np.ndarray:
        """(n,4) u64 view, no copy (core/types.rs:148-152)."""
        a = np.asarray(scalars)
        if a.dtype != np.uint64:
            a = np.frombuffer(a, dtype=np.uint64) if isinstance(scalars, (bytes, bytearray)) else a.view(np.uint64)
        return a.reshape(-1, 4)

    @staticmethod
    def g1_slice_as_icicle(points) -> np.ndarray:
        a = np.asarray(points)
        if a.dtype != np.uint64:
            a = a.view(np.uint64)
        return a.reshape(-1, 12)

    @staticmethod
    def g2_slice_as_icicle(points) -> np.ndarray:
        a = np.asarray(points)
        if a.dtype != np.uint64:
            a = a.view(np.uint64)
        return a.reshape(-1, 24)

    @staticmethod
    def _int(limbs) -> int:
        return sum(int(x) << (64 * i) for i, x in enumerate(limbs))

    @classmethod
    def icicle_to_g1_projective(cls, res: np.ndarray):
        """ICICLE G1 projective (standard form) -> affine (x, y) ints, or None for the identity
        (core/types.rs:353-368 does to_affine + from_xy)."""
        r = np.asarray(res, dtype=np.uint64).reshape(18)
        z = cls._int(r[12:18])
        if z == 0:
            return None
        assert z == 1, "backend returns normalised (x, y, 1)"
        return cls._int(r[0:6]), cls._int(r[6:12])

    @classmethod
    def icicle_to_g2_projective(cls, res: np.ndarray):
        r = np.asarray(res, dtype=np.uint64).reshape(36)
        z = (cls._int(r[24:30]), cls._int(r[30:36]))
        if z == (0, 0):
            return None
        return ((cls._int(r[0:6]), cls._int(r[6:12])), (cls._int(r[12:18]), cls._int(r[18:24])))
