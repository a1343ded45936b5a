#!/usr/bin/env python3
"""Does the level-0 gather of the affine pre-reduction get cheaper when the bases fit the L2?  (dev tool)
G1 MSM at c = 16 over 2^18..2^24 distinct resident bases; prints the level-0 forward / backward kernel time per pair sum
(b381_msm_last_level0_ms; n * 16 / 2 pair sums at level 0)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench as B
from midnight_bls12_381_cuda_b200 import _lib as L

os.environ["B381_MSM_TIMING"] = "1"
lib = L.lib()
g = np.array(B.G1_GEN_MONT, dtype=np.uint64)
for logn in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "18,19,20,21,22,24").split(",")]:
    n = 1 << logn
    bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(n), L.ptr(bases), None), "series")
    sc = B.canonical_fr(torch, n, 0xB12381)
    cfg = lib.b381_default_msm_config()
    cfg.are_scalars_on_device = cfg.are_points_on_device = True
    cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
    cfg.c = 16
    os.environ["B381_MSM_LEVELS"] = "2"
    res = np.zeros(18, dtype=np.uint64)
    f, b = C.c_float(), C.c_float()
    bf = bb = 1e9
    for it in range(4):
        L.check(lib.b381_g1_msm(L.ptr(sc), L.ptr(bases), n, C.byref(cfg), L.ptr(res)), "msm")
        torch.cuda.synchronize()
        if lib.b381_msm_last_level0_ms(C.byref(f), C.byref(b)) and it:
            bf, bb = min(bf, f.value), min(bb, b.value)
    pairs = n * 16 / 2
    print(f"2^{logn} ({n * 96 / 1e6:7.1f} MB of bases): level-0 fwd {bf:7.3f} ms = {bf * 1e9 / pairs:6.1f} ps/pair, "
          f"bwd {bb:7.3f} ms = {bb * 1e9 / pairs:6.1f} ps/pair", flush=True)
    del bases, sc
