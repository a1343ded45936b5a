#!/usr/bin/env python3
"""Affine pre-reduction sweep (dev tool): G1 MSM correctness + phase timings for B381_MSM_LEVELS in a list.
usage: gpu_check3.py <logn,logn,...> <levels,levels,...|d>   ('d' = library default heuristic)"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv, argv = sys.argv[:1], sys.argv
import numpy as np
import torch

import tools.gpu_check1 as G
from oracle import pyref as P

lib = G.lib


def timings():
    buf = (C.c_float * 12)()
    k = lib.b381_msm_last_timings(buf, 12)
    return [round(buf[i], 3) for i in range(k)]


def level0():
    f, b = C.c_float(), C.c_float()
    if lib.b381_msm_last_level0_ms(C.byref(f), C.byref(b)) == 1:
        return f"L0 fwd {f.value:.2f} bwd {b.value:.2f}"
    return ""


def main():
    sizes = [int(x) for x in argv[1].split(",")] if len(argv) > 1 else [20, 22]
    lv = argv[2].split(",") if len(argv) > 2 else ["0", "d"]
    m = 4096
    pts, ks = G.g1_pool(m, seed=21)
    pool = np.frombuffer(b"".join(P.g1_affine_mont_bytes(p) for p in pts), dtype=np.uint8).reshape(m, 96)
    os.environ["B381_MSM_TIMING"] = "1"
    for logn in sizes:
        n = 1 << logn
        dev = torch.from_numpy(np.tile(pool, (n // m, 1))).cuda()
        sc = G.rand_scalars_np(n, logn)
        sums = G.class_sums(sc, m)
        exp = P.g1_result_std_bytes(P.g1_mul(sum(s * k for s, k in zip(sums, ks)) % P.R_MOD, P.G1_GEN))
        for l in lv:
            if l == "d":
                os.environ.pop("B381_MSM_LEVELS", None)
            else:
                os.environ["B381_MSM_LEVELS"] = l
            got, dt = G.run_msm("g1", sc, dev, n)
            got, dt = G.run_msm("g1", sc, dev, n)
            print(f"g1 msm 2^{logn} levels={l}: {'OK' if got == exp else 'MISMATCH'} wall {dt*1e3:.2f} ms  "
                  f"phases(ms) {timings()} {level0()} -> {n/dt:.3e} pts/s", flush=True)
        del dev
        torch.cuda.empty_cache()
    print("ALL DONE")


if __name__ == "__main__":
    main()
