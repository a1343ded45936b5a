"""B200-native BLS12-381 proving backend (MSM G1/G2, Fr NTT, Fr vecops) -- Python host layer.

The product is the CUDA library (csrc/, C ABI in include/b381.h); this package mirrors the
reference's Rust `core/` API over that ABI via ctypes.  No CPU fallback exists.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
