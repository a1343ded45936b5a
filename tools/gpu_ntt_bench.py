#!/usr/bin/env python3
"""Fr NTT timing (dev tool): forward kNN in place, device resident, CUDA events, best and mean of `reps` after warm-up.
usage: gpu_ntt_bench.py <logn[:batch],...> [reps]   (B381_NTT_SHAPE / B381_NTT_GENERIC select kernel variants)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench as B
import midnight_bls12_381_cuda_b200 as M
from midnight_bls12_381_cuda_b200 import _lib as L

lib = L.lib()


def main():
    specs = [s.split(":") for s in sys.argv[1].split(",")]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    ctx = M.GpuNttContext(26)
    v, ms = C.c_double(), C.c_float()
    L.check(lib.b381_bench_field_mul(1, 1000, C.byref(v), C.byref(ms)), "fr probe")
    fr_rate = v.value
    print(f"shape={os.environ.get('B381_NTT_SHAPE', '0')} generic={os.environ.get('B381_NTT_GENERIC', '0')} fr_mul {fr_rate:.3e}/s")
    for sp in specs:
        logn, batch = int(sp[0]), int(sp[1]) if len(sp) > 1 else 1
        n = 1 << logn
        x = B.canonical_fr(torch, n * batch, 7)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for direction in (0, 1):
            ts = []
            for it in range(reps + 3):
                e0.record()
                ctx.ntt_on_device(x.data_ptr(), direction, size=n, batch=batch)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1))
            floor = 0.5 * n * batch * logn / fr_rate * 1e3
            best, mean = min(ts), sum(ts) / len(ts)
            print(f"  2^{logn} x{batch} dir={direction}: best {best:.3f} ms mean {mean:.3f} ms  {n * batch / (mean * 1e-3):.3e} elem/s  "
                  f"fr_mul floor {floor:.3f} ms ({floor / mean:.2f})", flush=True)


if __name__ == "__main__":
    main()
