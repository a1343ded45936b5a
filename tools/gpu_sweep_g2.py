#!/usr/bin/env python3
"""Window / level sweep of the G2 MSM (dev tool).  usage: gpu_sweep_g2.py <logn> <c,c,...> [levels,...|d]
bases (1+i)G2 laid down by b381_g2_point_series, uniform canonical Montgomery scalars; every configuration must return
the same bytes."""
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
logn = int(sys.argv[1])
cs = [int(x) for x in sys.argv[2].split(",")]
lvs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["d"]
n = 1 << logn
g2g = np.array(B.G2_GEN_MONT, dtype=np.uint64)
bases = torch.empty((n, 24), dtype=torch.int64, device="cuda")
L.check(lib.b381_g2_point_series(L.ptr(g2g), L.ptr(g2g), C.c_uint64(n), L.ptr(bases), None), "g2 series")
sc = B.canonical_fr(torch, n, 0xB12381)
os.environ["B381_MSM_TIMING"] = "1"
cfg = lib.b381_default_msm_config()
cfg.are_scalars_on_device = cfg.are_points_on_device = True
cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
ref = None
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for c in cs:
    for lv in lvs:
        if lv == "d":
            os.environ.pop("B381_MSM_LEVELS", None)
        else:
            os.environ["B381_MSM_LEVELS"] = lv
        cfg.c = c
        res = np.zeros(36, dtype=np.uint64)
        best = 1e9
        for it in range(3):
            ev0.record()
            L.check(lib.b381_g2_msm(L.ptr(sc), L.ptr(bases), n, C.byref(cfg), L.ptr(res)), "g2 msm")
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
        ok = "same" if res.tobytes() == ref else "MISMATCH"
        print(f"G2 2^{logn} c={info[0]} W={info[1]} levels={info[2]}: {best:.2f} ms {ok} phases {[round(buf[i], 2) for i in range(k)]}", flush=True)
print("SWEEP DONE")
