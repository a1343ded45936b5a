#!/usr/bin/env python3
"""One short invocation of a hot path, as the target of an `ncu` capture (dev tool).
usage: gpu_profile_target.py msm24|msm21|g2_20|ntt24|batch12   -- one call after one warm-up call."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench as B
import midnight_bls12_381_cuda_b200 as M
from midnight_bls12_381_cuda_b200 import _lib as L

lib = L.lib()
what = sys.argv[1]
g = np.array(B.G1_GEN_MONT, dtype=np.uint64)


def msm_g1(logn, batch=1):
    n = 1 << logn
    bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(n), L.ptr(bases), None), "series")
    sc = B.canonical_fr(torch, n * batch, 0xB12381)
    cfg = lib.b381_default_msm_config()
    cfg.are_scalars_on_device = cfg.are_points_on_device = True
    cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
    cfg.batch_size = batch
    res = np.zeros((batch, 18), dtype=np.uint64)
    for _ in range(2):
        L.check(lib.b381_g1_msm(L.ptr(sc), L.ptr(bases), n, C.byref(cfg), L.ptr(res)), "msm")
    torch.cuda.synchronize()


if what == "msm24":
    msm_g1(24)
elif what == "msm21":
    msm_g1(21)
elif what == "batch12":
    msm_g1(12, 16)
elif what == "g2_20":
    n = 1 << 20
    g2g = np.array(B.G2_GEN_MONT, dtype=np.uint64)
    bases = torch.empty((n, 24), dtype=torch.int64, device="cuda")
    L.check(lib.b381_g2_point_series(L.ptr(g2g), L.ptr(g2g), C.c_uint64(n), L.ptr(bases), None), "g2 series")
    sc = B.canonical_fr(torch, n, 0xB12381)
    cfg = lib.b381_default_msm_config()
    cfg.are_scalars_on_device = cfg.are_points_on_device = True
    cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
    res = np.zeros(36, dtype=np.uint64)
    for _ in range(2):
        L.check(lib.b381_g2_msm(L.ptr(sc), L.ptr(bases), n, C.byref(cfg), L.ptr(res)), "g2 msm")
    torch.cuda.synchronize()
elif what == "ntt24":
    ctx = M.GpuNttContext(24)
    vec = B.canonical_fr(torch, 1 << 24, 7)
    for _ in range(2):
        ctx.ntt_on_device(vec.data_ptr(), 0, size=1 << 24)
    torch.cuda.synchronize()
else:
    raise SystemExit("unknown target " + what)
print("target", what, "done")
