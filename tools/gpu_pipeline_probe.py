#!/usr/bin/env python3
"""Async MSM handles, two in flight: per-commit completion intervals (dev tool)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench as B
import midnight_bls12_381_cuda_b200 as M
from midnight_bls12_381_cuda_b200 import _lib as L
from midnight_bls12_381_cuda_b200.stream import DeviceVec

lib = L.lib()
n = 1 << 24
g = np.array(B.G1_GEN_MONT, dtype=np.uint64)
bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(n), L.ptr(bases), None), "series")
sc = B.canonical_fr(torch, n, 0xB12381)
sc_host = torch.empty(sc.shape, dtype=sc.dtype).pin_memory()
sc_host.copy_(sc)
torch.cuda.synchronize()
ctx = M.GpuMsmContext()
dev_bases = M.msm.PrecomputedBases(DeviceVec.borrow(bases.data_ptr(), n, 96), n)
sc_np = sc_host.numpy().view(np.uint64)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pending, stamps = [], []
t0 = time.perf_counter()
for i in range(14):
    ta = time.perf_counter()
    pending.append(ctx.msm_with_device_bases_async(sc_np, dev_bases))
    tb = time.perf_counter()
    if len(pending) == depth:
        pending.pop(0).wait()
        stamps.append((time.perf_counter() - t0, tb - ta))
while pending:
    pending.pop(0).wait()
    stamps.append((time.perf_counter() - t0, 0.0))
prev = 0.0
for t, launch in stamps:
    print(f"done at {t * 1e3:8.1f} ms  (+{(t - prev) * 1e3:6.1f})  launch call took {launch * 1e3:6.1f} ms")
    prev = t
free, total = torch.cuda.mem_get_info()
print("device memory in use (GB):", (total - free) / 1e9)
