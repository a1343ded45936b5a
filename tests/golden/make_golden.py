#!/usr/bin/env python3
"""Regenerates tests/golden/golden.json from the big-integer oracle (oracle/pyref.py).

Inputs follow the reference's own tests wherever it fixes them:
  * G1 MSM n=8, std::mt19937_64(12345), bases 2^i*G      bls12-381/tests/test_msm_security.cu:410-505
  * G1 MSM n=1024 seed 54321 / n=256 seed 99999           :634, :683
  * sum_{i=1..64} i*G = 2080*G, 5*G                       core/msm.rs:1667-1694
  * 1*G = G, 0*G = O, all-ones = sum of bases             test_msm_security.cu:908-941
  * NTT of 1..n at k=10                                    tests/ntt_fft_comparison.rs:15-19
  * NTT(delta_0) = (1,...,1)                               core/ntt.rs:2059-2073
  * coset generator 7                                      core/ntt.rs:2235
The reference holds NO absolute output vectors for MSM/NTT ("parity unpinned" for NTT
outputs, GPU-vs-BLST equality only in the consumer repo), so expected outputs here come from
pyref -- an implementation that shares no code with oracle.c or the CUDA path.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyref as P  # noqa: E402
from vectors import Mt19937_64, random_fr_integer  # noqa: E402


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def pow2_bases_g1(n):
    out, cur = [], P.G1_GEN
    for _ in range(n):
        out.append(cur)
        cur = P.g1_add(cur, cur)
    return out


def main():
    G = {"g1_msm": [], "g2_msm": [], "ntt": [], "vecops": [], "field": {}}
    # ---- field KATs
    G["field"]["fr_one_mont"] = hex(P.FR_R)
    G["field"]["fq_one_mont"] = hex(P.FQ_R)
    G["field"]["fr_root_of_unity_mont"] = hex(P.fr_to_mont(P.FR_ROOT_OF_UNITY))
    G["field"]["g1_gen_mont"] = P.g1_affine_mont_bytes(P.G1_GEN).hex()
    G["field"]["g2_gen_mont"] = P.g2_affine_mont_bytes(P.G2_GEN).hex()

    # ---- G1 MSM, reference-shaped cases (bases 2^i G, scalars integer form)
    for name, seed, n in (("ref_vs_scalar_mul_n8", 12345, 8), ("ref_n256_seed99999", 99999, 256),
                          ("ref_n1024_seed54321", 54321, 1024)):
        rng = Mt19937_64(seed)
        sc = [random_fr_integer(rng) for _ in range(n)]
        # sum s_i 2^i G
        k = sum(s << i for i, s in enumerate(sc)) % P.R_MOD
        G["g1_msm"].append({"name": name, "n": n, "bases": "pow2", "seed": seed,
                            "scalars_head": [hex(s) for s in sc[:4]],
                            "result": P.g1_result_std_bytes(P.g1_mul(k, P.G1_GEN)).hex()})
    G["g1_msm"].append({"name": "sum_i_times_G_64", "n": 64, "bases": "gen", "scalars": "iota1",
                        "result": P.g1_result_std_bytes(P.g1_mul(2080, P.G1_GEN)).hex()})
    G["g1_msm"].append({"name": "five_times_G", "n": 1, "bases": "gen", "scalars": [hex(5)],
                        "result": P.g1_result_std_bytes(P.g1_mul(5, P.G1_GEN)).hex()})
    G["g1_msm"].append({"name": "one_times_G", "n": 1, "bases": "gen", "scalars": [hex(1)],
                        "result": P.g1_result_std_bytes(P.G1_GEN).hex()})
    G["g1_msm"].append({"name": "zero_times_G", "n": 1, "bases": "gen", "scalars": [hex(0)],
                        "result": P.g1_result_std_bytes(None).hex()})
    G["g1_msm"].append({"name": "all_ones_pow2_16", "n": 16, "bases": "pow2", "scalars": "ones",
                        "result": P.g1_result_std_bytes(P.g1_mul((1 << 16) - 1, P.G1_GEN)).hex()})
    G["g1_msm"].append({"name": "r_minus_1_pow2_4", "n": 4, "bases": "pow2", "scalars": [hex(P.R_MOD - 1)] * 4,
                        "result": P.g1_result_std_bytes(P.g1_mul((P.R_MOD - 1) * 15, P.G1_GEN)).hex()})
    # explicit small random case with every byte spelled out
    rng = P.SplitMix64(0xB12381)
    ks = [rng.fr() for _ in range(6)]
    pts = [P.g1_mul(k, P.G1_GEN) for k in ks]
    pts[4] = None                                    # infinity base
    sc = [rng.fr() for _ in range(6)]
    exp = P.g1_msm(sc, pts)
    G["g1_msm"].append({"name": "explicit_n6_with_infinity", "n": 6,
                        "bases_mont_hex": [P.g1_affine_mont_bytes(p).hex() for p in pts],
                        "scalars": [hex(s) for s in sc], "result": P.g1_result_std_bytes(exp).hex()})
    # ---- G2
    ks = [rng.fr() for _ in range(5)]
    pts2 = [P.g2_mul(k, P.G2_GEN) for k in ks]
    sc = [rng.fr() for _ in range(5)]
    G["g2_msm"].append({"name": "explicit_n5", "n": 5, "bases_mont_hex": [P.g2_affine_mont_bytes(p).hex() for p in pts2],
                        "scalars": [hex(s) for s in sc], "result": P.g2_result_std_bytes(P.g2_msm(sc, pts2)).hex()})
    G["g2_msm"].append({"name": "sum_i_times_G2_32", "n": 32, "bases": "gen", "scalars": "iota1",
                        "result": P.g2_result_std_bytes(P.g2_mul(32 * 33 // 2, P.G2_GEN)).hex()})
    G["g2_msm"].append({"name": "zero_times_G2", "n": 1, "bases": "gen", "scalars": [hex(0)],
                        "result": P.g2_result_std_bytes(None).hex()})
    # ---- NTT (values are canonical integers; files store Montgomery bytes)
    def enc(vals):
        return b"".join(P.fr_bytes(P.fr_to_mont(v)) for v in vals)
    iota = list(range(1, 1025))
    y = P.ntt(iota)
    G["ntt"].append({"name": "iota_k10_forward_NN", "log_n": 10, "input": "iota1", "inverse": False, "ordering": "NN",
                     "coset": None, "sha256": sha(enc(y)), "head": [hex(P.fr_to_mont(v)) for v in y[:4]]})
    delta = [1] + [0] * 63
    G["ntt"].append({"name": "delta_k6", "log_n": 6, "input": "delta0", "inverse": False, "ordering": "NN", "coset": None,
                     "sha256": sha(enc(P.ntt(delta))), "head": [hex(P.fr_to_mont(1))] * 4})
    rng = P.SplitMix64(0xB12381_2020)
    for log_n in (4, 11, 12):
        vec = [rng.fr() for _ in range(1 << log_n)]
        for inverse in (False, True):
            for ordering in ("NN", "NR", "RN", "RR"):
                for g in (None, 7):
                    nat = P.apply_ordering(vec, ordering, "in")
                    out = P.coset_ntt(nat, g, inverse) if g else P.ntt(nat, inverse=inverse)
                    out = P.apply_ordering(out, ordering, "out")
                    G["ntt"].append({"name": f"rand_k{log_n}_{'inv' if inverse else 'fwd'}_{ordering}_{'coset7' if g else 'plain'}",
                                     "log_n": log_n, "input": f"splitmix:{0xB12381_2020}:{log_n}", "inverse": inverse,
                                     "ordering": ordering, "coset": g, "sha256": sha(enc(out)),
                                     "head": [hex(P.fr_to_mont(v)) for v in out[:2]]})
    # ---- vecops
    rng = P.SplitMix64(0xB12381_77)
    a = [rng.fr() for _ in range(8)] + [0, P.R_MOD - 1]
    b = [rng.fr() for _ in range(8)] + [0, P.R_MOD - 1]
    G["vecops"].append({"a": [hex(P.fr_to_mont(x)) for x in a], "b": [hex(P.fr_to_mont(x)) for x in b],
                        "add": [hex(P.fr_to_mont((x + y) % P.R_MOD)) for x, y in zip(a, b)],
                        "sub": [hex(P.fr_to_mont((x - y) % P.R_MOD)) for x, y in zip(a, b)],
                        "mul": [hex(P.fr_to_mont((x * y) % P.R_MOD)) for x, y in zip(a, b)],
                        "scalar_mul": [hex(P.fr_to_mont((a[0] * y) % P.R_MOD)) for y in b],
                        "scalar_add": [hex(P.fr_to_mont((a[0] + y) % P.R_MOD)) for y in b]})
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(G, f, indent=1)
    print("wrote golden.json:", {k: len(v) for k, v in G.items()})


if __name__ == "__main__":
    main()
