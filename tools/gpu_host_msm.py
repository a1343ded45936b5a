#!/usr/bin/env python3
"""One MSM configuration timed through the C ABI, resident or pinned-host scalars (dev tool).
usage: gpu_host_msm.py <g1|g2> <logn> <label,label,...> [host]  -- one timing line per label (labels only name the
lines: vary the library's environment knobs from the calling shell); G1 results are checked against the discrete-log
identity by the CPU oracle."""
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
grp = sys.argv[1]
logn = int(sys.argv[2])
ks = [x for x in sys.argv[3].split(",")]
host = len(sys.argv) > 4 and sys.argv[4] == "host"
n = 1 << logn
g1 = grp == "g1"
gen = np.array(B.G1_GEN_MONT if g1 else B.G2_GEN_MONT, dtype=np.uint64)
bases = torch.empty((n, 12 if g1 else 24), dtype=torch.int64, device="cuda")
series = lib.b381_g1_point_series if g1 else lib.b381_g2_point_series
L.check(series(L.ptr(gen), L.ptr(gen), C.c_uint64(n), L.ptr(bases), None), "series")
sc = B.canonical_fr(torch, n, 0xB12381)
sc_in = sc
if host:
    sc_in = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    sc_in.copy_(sc)
os.environ["B381_MSM_TIMING"] = "1"
cfg = lib.b381_default_msm_config()
cfg.are_scalars_on_device, cfg.are_points_on_device = not host, True
cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
fn = lib.b381_g1_msm if g1 else lib.b381_g2_msm
ref = None
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in ks:
    res = np.zeros(18 if g1 else 36, dtype=np.uint64)
    best = 1e9
    for it in range(4):
        ev0.record()
        L.check(fn(L.ptr(sc_in), L.ptr(bases), n, C.byref(cfg), L.ptr(res)), "msm")
        ev1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, ev0.elapsed_time(ev1))
    buf = (C.c_float * 12)()
    cnt = lib.b381_msm_last_timings(buf, 12)
    info = (C.c_int * 4)()
    lib.b381_msm_last_info(info, 4)
    if ref is None:
        ref = res.tobytes()
        ok = "ref"
        if g1 and logn <= 24:
            from oracle import cref as O
            from oracle import pyref as Pr
            kk = np.zeros((n, 4), dtype=np.uint64)
            kk[:, 0] = np.arange(1, n + 1, dtype=np.uint64)
            dl = Pr.from_limbs(O.fr_dot(sc.cpu().numpy().view(np.uint64), kk, s_mont=True))
            ok = "ORACLE-OK" if ref == Pr.g1_result_std_bytes(Pr.g1_mul(dl, Pr.G1_GEN)) else "ORACLE-MISMATCH"
    else:
        ok = "same" if res.tobytes() == ref else "MISMATCH"
    print(f"{grp} 2^{logn}{' host' if host else ''} {k} c={info[0]} W={info[1]} levels={info[2]} launches={info[3]}: "
          f"{best:.2f} ms {ok} phases {[round(buf[i], 2) for i in range(cnt)]}", flush=True)
print("SWEEP DONE")
