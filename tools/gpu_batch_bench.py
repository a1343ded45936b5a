#!/usr/bin/env python3
"""Batched G1 MSM over shared resident bases: one folded pipeline run (batch_size = B) against B separate calls.
usage: gpu_batch_bench.py <logn,batch> ...   (dev tool; numbers go to profiles/)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench as B
from midnight_bls12_381_cuda_b200 import _lib as L

lib = L.lib()
g = np.array(B.G1_GEN_MONT, dtype=np.uint64)
for spec in sys.argv[1:]:
    logn, batch = (int(x) for x in spec.split(","))
    n = 1 << logn
    bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(n), L.ptr(bases), None), "series")
    sc = B.canonical_fr(torch, n * batch, 0xBA7C4)
    cfg = lib.b381_default_msm_config()
    cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
    cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
    res_b = torch.zeros((batch, 18), dtype=torch.int64, device="cuda")
    res_s = torch.zeros((batch, 18), dtype=torch.int64, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def folded():
        cfg.batch_size = batch
        L.check(lib.b381_g1_msm(L.ptr(sc), L.ptr(bases), n, C.byref(cfg), L.ptr(res_b)), "batch")

    def separate():
        cfg.batch_size = 1
        for b in range(batch):
            L.check(lib.b381_g1_msm(C.c_void_p(sc.data_ptr() + 32 * n * b), L.ptr(bases), n, C.byref(cfg),
                                    C.c_void_p(res_s.data_ptr() + 144 * b)), "single")

    out = {}
    for name, fn in (("folded", folded), ("separate", separate)):
        best = 1e9
        for it in range(4):
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            if it:
                best = min(best, ev0.elapsed_time(ev1))
        out[name] = best
    same = bool((res_b == res_s).all())
    print(f"2^{logn} x {batch}: folded {out['folded']:.3f} ms, {batch} separate calls {out['separate']:.3f} ms "
          f"({out['separate'] / out['folded']:.2f}x), results {'equal' if same else 'DIFFER'}", flush=True)
    del bases, sc
