#!/usr/bin/env python3
"""Run the roofline-denominator probes (IMAD.WIDE issue rate, Fq/Fr Montgomery mul rate)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midnight_bls12_381_cuda_b200 import _lib as L
lib = L.lib()
ms = C.c_float(); v = C.c_double()
it = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L.check(lib.b381_bench_imad_peak(it * 4, C.byref(v), C.byref(ms)), "imad"); print(f"imad.wide {v.value:.4e} MAD/s {ms.value:.3f} ms")
for f, name, mads in ((0, "fq", 300), (1, "fr", 136)):
    L.check(lib.b381_bench_field_mul(f, it, C.byref(v), C.byref(ms)), "mul")
    print(f"{name} mul {v.value:.4e} mul/s = {v.value*mads:.4e} MAD/s  {ms.value:.3f} ms")
