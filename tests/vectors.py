"""Deterministic test inputs shared by the CPU and GPU suites.

`Mt19937_64` reproduces std::mt19937_64 so that the reference's own MSM test inputs
(bls12-381/tests/test_msm_security.cu:410-505 seed 12345, :634 seed 54321, :683 seed 99999, drawn
with random_fr_integer, tests/security_audit_tests.cuh:400-416) can be regenerated bit for bit.
"""
from __future__ import annotations

import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
M64 = (1 << 64) - 1


class Mt19937_64:
    NN, MM = 312, 156

    def __init__(self, seed: int):
        mt = [0] * self.NN
        mt[0] = seed & M64
        for i in range(1, self.NN):
            mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & M64
        self.mt, self.i = mt, self.NN

    def __call__(self) -> int:
        if self.i >= self.NN:
            mt, NN, MM = self.mt, self.NN, self.MM
            for i in range(NN):
                x = (mt[i] & 0xFFFFFFFF80000000) | (mt[(i + 1) % NN] & 0x7FFFFFFF)
                mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (0xB5026F5AA96619E9 if x & 1 else 0)
            self.i = 0
        x = self.mt[self.i]
        self.i += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x & M64


def random_fr_integer(rng) -> int:
    """security_audit_tests.cuh:400-416"""
    while True:
        l = [rng() for _ in range(4)]
        l[3] &= 0x7FFFFFFFFFFFFFFF
        v = sum(x << (64 * i) for i, x in enumerate(l))
        if v < R_MOD:
            return v


def limbs(v: int, n: int):
    return [(v >> (64 * i)) & M64 for i in range(n)]


def fr_array(vals) -> np.ndarray:
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i] = limbs(v, 4)
    return a


def fr_ints(arr) -> list[int]:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(x) << (64 * i) for i, x in enumerate(row)) for row in arr]


# --------------------------------------------------------------------------- golden-case inputs
def load_golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")) as f:
        return json.load(f)


_pow2_cache = {}


def msm_case_inputs(case, group: str):
    """(scalars (n,4) integer form, bases (n,12k) Montgomery affine) of a golden MSM case,
    built with the big-integer oracle only."""
    from oracle import pyref as P
    n = case["n"]
    k = 1 if group == "g1" else 2
    enc = P.g1_affine_mont_bytes if k == 1 else P.g2_affine_mont_bytes
    gen = P.G1_GEN if k == 1 else P.G2_GEN
    add = P.g1_add if k == 1 else P.g2_add
    if "bases_mont_hex" in case:
        raw = b"".join(bytes.fromhex(h) for h in case["bases_mont_hex"])
    elif case["bases"] == "gen":
        raw = enc(gen) * n
    elif case["bases"] == "pow2":
        key = (group, n)
        if key not in _pow2_cache:
            out, cur = [], gen
            for _ in range(n):
                out.append(enc(cur))
                cur = add(cur, cur)
            _pow2_cache[key] = b"".join(out)
        raw = _pow2_cache[key]
    bases = np.frombuffer(raw, dtype=np.uint64).reshape(n, 12 * k).copy()
    sc = case.get("scalars")
    if "seed" in case:
        rng = Mt19937_64(case["seed"])
        vals = [random_fr_integer(rng) for _ in range(n)]
        assert [hex(v) for v in vals[:4]] == case["scalars_head"][: min(4, n)]
    elif sc == "iota1":
        vals = list(range(1, n + 1))
    elif sc == "ones":
        vals = [1] * n
    else:
        vals = [int(h, 16) for h in sc]
    return fr_array(vals), bases


def ntt_case_input(case):
    """canonical-integer input vector of a golden NTT case."""
    from oracle import pyref as P
    n = 1 << case["log_n"]
    if case["input"] == "iota1":
        return list(range(1, n + 1))
    if case["input"] == "delta0":
        return [1] + [0] * (n - 1)
    _, seed, _ = case["input"].split(":")
    rng = P.SplitMix64(int(seed))
    out = None
    for log_n in (4, 11, 12):             # make_golden.py draws the three sizes from one stream
        vec = [rng.fr() for _ in range(1 << log_n)]
        if log_n == case["log_n"]:
            out = vec
    return out


ORDERINGS = {"NN": 0, "NR": 1, "RN": 2, "RR": 3}
