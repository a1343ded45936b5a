#!/usr/bin/env python3
"""GPU bring-up 2: NTT correctness (vs big-int oracle) and timing (dev tool)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from midnight_bls12_381_cuda_b200 import _lib as L
from oracle import pyref as P

lib = L.lib()
ORD = ['NN', 'NR', 'RN', 'RR']

def fr_arr(vals):
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i] = P.to_limbs(v, 4)
    return a

def init(K, mont=False):
    w = P.fr_omega(K)
    root = fr_arr([P.fr_to_mont(w) if mont else w])
    cfg = L.NTTInitDomainConfig()
    L.check(lib.b381_ntt_init_domain(L.ptr(root), C.byref(cfg)), "init")

def ntt_dev(d_in, d_out, N, inverse, ordering=0, batch=1, g=None, columns=False):
    cfg = lib.b381_default_ntt_config()
    cfg.batch_size = batch
    cfg.ordering = ordering
    cfg.columns_batch = columns
    cfg.are_inputs_on_device = True
    cfg.are_outputs_on_device = True
    if g is not None:
        for i, l in enumerate(P.to_limbs(P.fr_to_mont(g), 4)):
            cfg.coset_gen.l[i] = l
    L.check(lib.b381_ntt(L.ptr(d_in), N, 1 if inverse else 0, C.byref(cfg), L.ptr(d_out)), "ntt")

def expect(vec, inverse, ordering, g):
    o = ORD[ordering]
    nat = P.apply_ordering(vec, o, 'in')
    y = P.coset_ntt(nat, g, inverse) if g else P.ntt(nat, inverse=inverse)
    return P.apply_ordering(y, o, 'out')

def small():
    rng = P.SplitMix64(3)
    for n in (1, 4, 10, 11, 12, 14):
        N = 1 << n
        vec = [rng.fr() for _ in range(N)]
        d = torch.from_numpy(fr_arr([P.fr_to_mont(v) for v in vec]).view(np.int64)).cuda()
        for inverse in (False, True):
            for ordering in range(4):
                for g in (None, 7):
                    out = torch.empty_like(d)
                    ntt_dev(d, out, N, inverse, ordering, g=g)
                    got = out.cpu().numpy().view(np.uint64)
                    exp = fr_arr([P.fr_to_mont(v) for v in expect(vec, inverse, ordering, g)])
                    assert (got == exp).all(), (n, inverse, ordering, g)
                    d2 = d.clone()
                    ntt_dev(d2, d2, N, inverse, ordering, g=g)     # in place
                    assert (d2.cpu().numpy().view(np.uint64) == exp).all(), ("inplace", n, inverse, ordering, g)
        print("ntt ok n =", n, flush=True)
    # k=10 input 1..n (tests/ntt_fft_comparison.rs:15-19), host buffers
    N = 1024
    vec = list(range(1, N + 1))
    h_in = fr_arr([P.fr_to_mont(v) for v in vec]); h_out = np.zeros_like(h_in)
    cfg = lib.b381_default_ntt_config()
    L.check(lib.b381_ntt(L.ptr(h_in), N, 0, C.byref(cfg), L.ptr(h_out)), "ntt host")
    assert (h_out == fr_arr([P.fr_to_mont(v) for v in P.ntt(vec)])).all()
    print("ntt 1..n host-buffer ok")

def large():
    g = np.random.default_rng(1)
    for n in (16, 20, 22, 24):
        N = 1 << n
        a = g.integers(0, 1 << 63, size=(N, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 62) - 1)
        d = torch.from_numpy(a.view(np.int64)).cuda()
        out = torch.empty_like(d); back = torch.empty_like(d)
        for ordering in (0, 1):
            ntt_dev(d, out, N, False, ordering); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ntt_dev(d, out, N, False, ordering)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"ntt 2^{n} {ORD[ordering]}: {ms:.3f} ms  {N/ms*1e3:.3e} elem/s  ({128*N/ms/1e6:.0f} GB/s alg.)", flush=True)
        # round trip NR -> RN
        ntt_dev(d, out, N, False, 1)
        ntt_dev(out, back, N, True, 2)
        assert torch.equal(back, d), "round trip"
        ntt_dev(d, out, N, False, 0)
        # Horner spot checks: y[i] = sum a[j] w^(ij)
        if n <= 20:
            w = P.fr_omega(n)
            vals = [P.fr_from_mont(P.from_limbs(r)) for r in a]  # treat bytes as Montgomery form
            o = out.cpu().numpy().view(np.uint64)
            for i in (0, 1, N // 2 + 3, N - 1):
                wi = pow(w, i, P.R_MOD); acc = 0
                for v in reversed(vals): acc = (acc * wi + v) % P.R_MOD
                assert P.from_limbs(o[i]) == P.fr_to_mont(acc), ("horner", n, i)
            print("  horner ok")
        print("  roundtrip ok")

if __name__ == "__main__":
    init(24)
    small()
    large()
    print("ALL DONE")
