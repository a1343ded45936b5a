#!/usr/bin/env python3
"""precompute_factor sweep of the G1 MSM on resident bases (dev tool; the numbers go to profiles/).
usage: gpu_sweep_factor.py <logn> <factor,factor,...>  -- bases (1+i)G, uniform Montgomery scalars; the table
out[i*f + k] = 2^(k*c*Wf) P_i is built by b381_g1_msm_precompute_bases (outside the timed region: SRS setup) and
every factor must return the bytes of factor 1."""
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


def main():
    logn = int(sys.argv[1])
    fs = [int(x) for x in sys.argv[2].split(",")]
    n = 1 << logn
    g = np.array(B.G1_GEN_MONT, dtype=np.uint64)
    bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(n), L.ptr(bases), None), "series")
    sc = B.canonical_fr(torch, n, 0xB12381)
    os.environ["B381_MSM_TIMING"] = "1"
    ref = None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for f in fs:
        cfg = lib.b381_default_msm_config()
        cfg.are_scalars_on_device = cfg.are_points_on_device = True
        cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
        cfg.c = 16
        cfg.precompute_factor = f
        table = bases
        if f > 1:
            table = torch.empty((n * f, 12), dtype=torch.int64, device="cuda")
            cfg.are_results_on_device = True
            ev0.record()
            L.check(lib.b381_g1_msm_precompute_bases(L.ptr(bases), n, C.byref(cfg), L.ptr(table)), "precompute")
            ev1.record()
            torch.cuda.synchronize()
            pre_ms = ev0.elapsed_time(ev1)
            cfg.are_results_on_device = False
        else:
            pre_ms = 0.0
        res = np.zeros(18, dtype=np.uint64)
        best = 1e9
        for it in range(4):
            ev0.record()
            L.check(lib.b381_g1_msm(L.ptr(sc), L.ptr(table), n, C.byref(cfg), L.ptr(res)), "msm")
            ev1.record()
            torch.cuda.synchronize()
            if it:
                best = min(best, ev0.elapsed_time(ev1))
        buf = (C.c_float * 12)()
        k = lib.b381_msm_last_timings(buf, 12)
        info = (C.c_int * 4)()
        lib.b381_msm_last_info(info, 4)
        if ref is None:
            ref = res.tobytes()
            ok = "ref"
        else:
            ok = "same" if res.tobytes() == ref else "MISMATCH"
        print(f"2^{logn} factor={f} W={info[1]} levels={info[2]}: {best:.2f} ms {ok} (table {pre_ms:.0f} ms) phases {[round(buf[i], 2) for i in range(k)]}", flush=True)
        del table
        torch.cuda.empty_cache()
    print("SWEEP DONE")


if __name__ == "__main__":
    main()
