"""CPU-only execution of the per-thread bodies that the CUDA kernels are made of
(csrc/msm_core.cuh, csrc/ntt_core.cuh, csrc/curve.cuh) via tests/host/*_host_sim.cpp, checked
against the big-integer oracle.  This is how kernel LOGIC is covered where no GPU exists; the
inline-PTX arithmetic underneath is covered by test_gen_field.py and, on the GPU, by -m gpu."""
import itertools
import os
import subprocess

import pytest

from oracle import pyref as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "midnight_bls12_381_cuda_b200", "csrc")
HOST = os.path.join(ROOT, "tests", "host")


@pytest.fixture(scope="module")
def sims(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostsim")
    bins = {}
    for name in ("msm_host_sim", "ntt_host_sim"):
        exe = str(out / name)
        subprocess.run(["g++", "-O2", "-std=c++17", f"-I{HOST}", f"-I{CSRC}", "-o", exe,
                        os.path.join(HOST, name + ".cpp")], check=True)
        bins[name] = exe
    bins["dir"] = str(out)
    return bins


def run_msm(sims, group, scalars, pts, c, K, L, mont=True, factor=1, levels=0, batch=1, shared=True, chunk_log=31,
            piece_chunks=0):
    """batch > 1: `scalars` holds batch * n values ([batch][n]); returns the list of results"""
    path = os.path.join(sims["dir"], "msm_in.bin")
    with open(path, "wb") as f:
        for s in scalars:
            f.write(P.fr_bytes(P.fr_to_mont(s) if mont else s))
        for pt in pts:
            f.write(P.g1_affine_mont_bytes(pt) if group == "g1" else P.g2_affine_mont_bytes(pt))
    r = subprocess.run([sims["msm_host_sim"], group, str(len(scalars) // batch), str(c), str(K), str(L), str(int(mont)), path,
                        str(factor), str(levels), str(batch), str(int(shared)), str(chunk_log),
                        str(piece_chunks)],
                       capture_output=True, text=True, check=True)
    out = [bytes.fromhex(x) for x in r.stdout.split()]
    return out[0] if batch == 1 else out


def test_msm_chunk_major_grouping(sims, tmp_path):
    """msm_core.cuh CHUNK-MAJOR GROUPING: entries grouped by (chunk of scalars, bucket slot), level 0 pairs inside runs
    and writes its sums bucket-major; same result as the plain grouping, G1 and G2, batch, folding, exceptional pairs
    (repeated bases, P - P, infinity), chunk sizes that do not divide n -- also once under ASan + UBSan."""
    exe = str(tmp_path / "msm_host_sim_asan")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-D_GLIBCXX_ASSERTIONS", f"-I{HOST}",
                    f"-I{CSRC}", "-o", exe, os.path.join(HOST, "msm_host_sim.cpp")], check=True)
    asan = dict(sims)
    asan["msm_host_sim"] = exe
    rng = P.SplitMix64(1919)
    for (group, n, batch, c, K, L, factor, levels, chunk_log, which) in [
            ("g1", 70, 1, 4, 3, 4, 1, 3, 4, sims), ("g1", 64, 1, 3, 50, 2, 1, 1, 3, sims), ("g1", 45, 3, 4, 3, 2, 2, 2, 5, sims),
            ("g1", 90, 1, 3, 4, 2, 1, 6, 4, asan), ("g2", 20, 2, 4, 3, 4, 1, 2, 3, asan)]:
        mul, gen = (P.g1_mul, P.G1_GEN) if group == "g1" else (P.g2_mul, P.G2_GEN)
        ks = [rng.fr() for _ in range(n)]
        pts = [mul(k, gen) for k in ks]
        pts[5], ks[5] = None, 0
        pts[7], ks[7] = pts[6], ks[6]                              # P + P inside a run
        pts[9], ks[9] = (P.g1_neg(pts[8]) if group == "g1" else P.g2_neg(pts[8])), P.R_MOD - ks[8]
        sc = [rng.fr() for _ in range(n * batch)]
        sc[1], sc[2], sc[7], sc[9] = 0, sc[3], sc[6], sc[8]
        got = run_msm(which, group, sc, pts, c, K, L, factor=factor, levels=levels, batch=batch, chunk_log=chunk_log)
        got = [got] if batch == 1 else got
        for b in range(batch):
            dl = sum(s * k for s, k in zip(sc[b * n:(b + 1) * n], ks)) % P.R_MOD
            exp = P.g1_result_std_bytes(mul(dl, gen)) if group == "g1" else P.g2_result_std_bytes(mul(dl, gen))
            assert got[b] == exp, (group, n, batch, chunk_log, b)


def test_msm_streamed_level0(sims, tmp_path):
    """Level 0 of a chunk-major run streamed piece by piece (msm_impl.cuh, plugin call with host scalars): per piece the
    half counts + a scan continued through the carried slot total, the forward pass of the threads that piece
    completes (pair_piece_owns) seeing only the offsets that exist so far; destinations in a second walk.  The host
    simulation checks that every thread runs exactly once and that the pieced scans equal the one-shot ones; the
    result must equal the discrete-log ground truth.  Zero scalars (all-trash pieces), one- and many-chunk pieces,
    a ragged last chunk; once under ASan + UBSan."""
    exe = str(tmp_path / "msm_host_sim_asan")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-D_GLIBCXX_ASSERTIONS", f"-I{HOST}",
                    f"-I{CSRC}", "-o", exe, os.path.join(HOST, "msm_host_sim.cpp")], check=True)
    asan = dict(sims)
    asan["msm_host_sim"] = exe
    rng = P.SplitMix64(2024)
    for (group, n, c, K, L, levels, chunk_log, piece_chunks, zeros, which) in [
            ("g1", 70, 4, 3, 4, 3, 3, 1, (), sims), ("g1", 70, 4, 3, 4, 2, 3, 2, (), sims), ("g1", 100, 3, 4, 2, 4, 4, 3, (), asan),
            ("g1", 64, 4, 3, 4, 2, 3, 2, range(16, 48), sims), ("g1", 40, 5, 3, 4, 1, 2, 3, range(0, 8), sims),
            ("g2", 24, 4, 3, 4, 2, 2, 2, (), asan)]:
        mul, gen = (P.g1_mul, P.G1_GEN) if group == "g1" else (P.g2_mul, P.G2_GEN)
        ks = [rng.fr() for _ in range(n)]
        pts = [mul(k, gen) for k in ks]
        pts[7], ks[7] = pts[6], ks[6]
        sc = [rng.fr() for _ in range(n)]
        sc[7] = sc[6]
        for i in zeros:
            sc[i] = 0
        got = run_msm(which, group, sc, pts, c, K, L, levels=levels, chunk_log=chunk_log, piece_chunks=piece_chunks)
        dl = sum(s * k for s, k in zip(sc, ks)) % P.R_MOD
        exp = P.g1_result_std_bytes(mul(dl, gen)) if group == "g1" else P.g2_result_std_bytes(mul(dl, gen))
        assert got == exp, (group, n, chunk_log, piece_chunks)


def test_msm_batch_folded_into_one_run(sims):
    """MSMConfig.batch_size: the MSMs of a batch share one pipeline run (bucket sets b*Wf ..), with shared bases and with
    per-MSM bases, with precomputed-bases folding and with affine levels"""
    rng = P.SplitMix64(4711)
    for (n, batch, shared, c, K, L, factor, levels) in [(9, 3, True, 4, 2, 2, 1, 0), (20, 4, False, 5, 3, 4, 1, 2),
                                                        (33, 2, True, 4, 3, 2, 2, 1), (16, 5, True, 3, 50, 2, 1, 3)]:
        np_ = n if shared else n * batch
        ks = [rng.fr() for _ in range(np_)]
        pts = [P.g1_mul(k, P.G1_GEN) for k in ks]
        sc = [rng.fr() for _ in range(n * batch)]
        sc[1], sc[n + 2] = 0, P.R_MOD - 1
        got = run_msm(sims, "g1", sc, pts, c, K, L, factor=factor, levels=levels, batch=batch, shared=shared)
        assert len(got) == batch
        for b in range(batch):
            kb = ks if shared else ks[b * n:(b + 1) * n]
            dl = sum(s * k for s, k in zip(sc[b * n:(b + 1) * n], kb)) % P.R_MOD
            assert got[b] == P.g1_result_std_bytes(P.g1_mul(dl, P.G1_GEN)), (n, batch, shared, b)


def test_msm_pipeline_g1(sims):
    rng = P.SplitMix64(42)
    for (n, c, K, L) in [(1, 4, 2, 2), (8, 4, 2, 4), (33, 5, 3, 4), (64, 8, 16, 16), (150, 7, 4, 8)]:
        ks = [rng.fr() for _ in range(n)]
        pts = [P.g1_mul(k, P.G1_GEN) for k in ks]
        sc = [rng.fr() for _ in range(n)]
        if n > 4:
            sc[1], sc[2], sc[3], pts[4] = 0, 1, P.R_MOD - 1, None
        exp = P.g1_mul(sum(s * k for s, k, p in zip(sc, ks, pts) if p is not None) % P.R_MOD, P.G1_GEN)
        assert run_msm(sims, "g1", sc, pts, c, K, L) == P.g1_result_std_bytes(exp), (n, c)


def test_msm_affine_prereduction_levels(sims):
    """csrc/msm_batch.cuh: pairwise affine levels in front of the task/accumulate path, G1 and G2, including the
    exceptional pairs (P+P, P-P, infinity operands, odd leftovers, empty buckets, slots past the end)."""
    rng = P.SplitMix64(77)
    for (n, c, K, L, levels) in [(1, 4, 2, 2, 1), (9, 3, 2, 2, 2), (64, 4, 3, 4, 1), (64, 4, 3, 4, 3), (150, 5, 4, 8, 2), (150, 3, 50, 2, 6)]:
        ks = [rng.fr() for _ in range(n)]
        pts = [P.g1_mul(k, P.G1_GEN) for k in ks]
        sc = [rng.fr() for _ in range(n)]
        if n > 4:
            sc[1], sc[2], sc[3], pts[4] = 0, 1, P.R_MOD - 1, None
        exp = P.g1_mul(sum(s * k for s, k, p in zip(sc, ks, pts) if p is not None) % P.R_MOD, P.G1_GEN)
        assert run_msm(sims, "g1", sc, pts, c, K, L, levels=levels) == P.g1_result_std_bytes(exp), (n, c, levels)
    for levels in (1, 2, 5):
        # identical bases: every pair is a doubling; then P + (-P): every pair cancels; infinity operands
        assert run_msm(sims, "g1", [5] * 40, [P.G1_GEN] * 40, 4, 3, 2, levels=levels) == P.g1_result_std_bytes(P.g1_mul(200, P.G1_GEN))
        pts = [P.G1_GEN, P.g1_neg(P.G1_GEN)] * 10
        assert run_msm(sims, "g1", [7] * 20, pts, 4, 3, 2, levels=levels) == P.g1_result_std_bytes(None)
        pts = [P.G1_GEN, None, None, P.G1_GEN, None] * 4
        assert run_msm(sims, "g1", [3] * 20, pts, 4, 3, 2, levels=levels) == P.g1_result_std_bytes(P.g1_mul(24, P.G1_GEN))
    n = 14
    ks = [rng.fr() for _ in range(n)]
    sc = [rng.fr() for _ in range(n)]
    sc[3] = sc[4] = sc[5]
    ks[4] = ks[3]                         # equal points in one bucket -> Fq2 doubling
    ks[5] = P.R_MOD - ks[3]               # and a cancellation
    pts2 = [P.g2_mul(k, P.G2_GEN) for k in ks]
    dl = sum(s * k for s, k in zip(sc, ks)) % P.R_MOD
    for levels in (1, 3):
        assert run_msm(sims, "g2", sc, pts2, 4, 3, 4, levels=levels) == P.g2_result_std_bytes(P.g2_mul(dl, P.G2_GEN))


def test_msm_bodies_under_address_sanitizer(sims, tmp_path):
    """compute-sanitizer is not available on the GPU pool, so the per-thread bodies (slot walk, slot-major scratch
    indexing, batched inversion, task splitting) are run once under ASan + UBSan with exactly-sized host arrays."""
    exe = str(tmp_path / "msm_host_sim_asan")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-D_GLIBCXX_ASSERTIONS", f"-I{HOST}",
                    f"-I{CSRC}", "-o", exe, os.path.join(HOST, "msm_host_sim.cpp")], check=True)
    asan = dict(sims)
    asan["msm_host_sim"] = exe
    rng = P.SplitMix64(2024)
    for (group, n, c, K, L, levels) in [("g1", 70, 4, 3, 4, 3), ("g1", 90, 3, 50, 2, 6), ("g2", 14, 4, 3, 4, 2)]:
        ks = [rng.fr() for _ in range(n)]
        sc = [rng.fr() for _ in range(n)]
        sc[1], sc[2] = 0, sc[3]
        mul, gen = (P.g1_mul, P.G1_GEN) if group == "g1" else (P.g2_mul, P.G2_GEN)
        pts = [mul(k, gen) for k in ks]
        pts[5] = None
        dl = sum(s * k for s, k, p in zip(sc, ks, pts) if p is not None) % P.R_MOD
        exp = P.g1_result_std_bytes(mul(dl, gen)) if group == "g1" else P.g2_result_std_bytes(mul(dl, gen))
        assert run_msm(asan, group, sc, pts, c, K, L, levels=levels) == exp, (group, n, levels)


def test_msm_pipeline_exceptional_cases(sims):
    # identical bases force the doubling branch of the mixed addition (benches/gpu_msm.rs:30-32)
    assert run_msm(sims, "g1", [5] * 40, [P.G1_GEN] * 40, 4, 3, 2) == P.g1_result_std_bytes(P.g1_mul(200, P.G1_GEN))
    # P + (-P) inside one bucket
    pts = [P.G1_GEN, P.g1_neg(P.G1_GEN)] * 10
    assert run_msm(sims, "g1", [7] * 20, pts, 4, 3, 2) == P.g1_result_std_bytes(None)
    # sum i*G = 2080 G with integer-form scalars (core/msm.rs:1681-1694)
    assert run_msm(sims, "g1", list(range(1, 65)), [P.G1_GEN] * 64, 6, 4, 4, mont=False) == \
        P.g1_result_std_bytes(P.g1_mul(2080, P.G1_GEN))


def test_msm_pipeline_g2_and_precompute(sims):
    rng = P.SplitMix64(43)
    n = 12
    ks = [rng.fr() for _ in range(n)]
    sc = [rng.fr() for _ in range(n)]
    pts2 = [P.g2_mul(k, P.G2_GEN) for k in ks]
    dl = sum(s * k for s, k in zip(sc, ks)) % P.R_MOD
    assert run_msm(sims, "g2", sc, pts2, 5, 3, 4) == P.g2_result_std_bytes(P.g2_mul(dl, P.G2_GEN))
    pts1 = [P.g1_mul(k, P.G1_GEN) for k in ks]
    for factor in (2, 5):
        assert run_msm(sims, "g1", sc, pts1, 6, 4, 4, factor=factor) == P.g1_result_std_bytes(P.g1_mul(dl, P.G1_GEN))


ORD = ["NN", "NR", "RN", "RR"]


def run_ntt(sims, n, batch, inverse, ordering, columns, g, inplace, vecs, tile_log=11):
    N = 1 << n
    flat = [0] * (N * batch)
    for b in range(batch):
        for i in range(N):
            flat[(i * batch + b) if columns else (b * N + i)] = vecs[b][i]
    path = os.path.join(sims["dir"], "ntt_in.bin")
    with open(path, "wb") as f:
        f.write(P.fr_bytes(P.fr_to_mont(P.fr_omega(max(n, 1)))))
        f.write(P.fr_bytes(P.fr_to_mont(g if g else 1)))
        for v in flat:
            f.write(P.fr_bytes(P.fr_to_mont(v)))
    r = subprocess.run([sims["ntt_host_sim"], str(n), str(batch), str(int(inverse)), str(ordering), str(int(columns)),
                        str(int(bool(g))), str(int(inplace)), path, str(tile_log)], capture_output=True, text=True, check=True)
    raw = bytes.fromhex(r.stdout.strip())
    flat = [P.fr_from_mont(int.from_bytes(raw[32 * i:32 * i + 32], "little")) for i in range(N * batch)]
    return [[flat[(i * batch + b) if columns else (b * N + i)] for i in range(N)] for b in range(batch)]


def expect_ntt(vec, inverse, ordering, g):
    o = ORD[ordering]
    nat = P.apply_ordering(vec, o, "in")
    y = P.coset_ntt(nat, g, inverse) if g else P.ntt(nat, inverse=inverse)
    return P.apply_ordering(y, o, "out")


@pytest.mark.parametrize("n", [0, 1, 3, 6, 11, 12, 13])
def test_ntt_passes(sims, n):
    rng = P.SplitMix64(99 + n)
    for batch, columns in ((1, False), (3, False), (2, True)):
        if n >= 12 and batch == 3:
            continue
        vecs = [[rng.fr() for _ in range(1 << n)] for _ in range(batch)]
        combos = list(itertools.product((False, True), range(4), (0, 7), (False, True)))
        if n >= 11:
            combos = combos[::7]
        for inverse, ordering, g, inplace in combos:
            got = run_ntt(sims, n, batch, inverse, ordering, columns, g, inplace, vecs)
            for b in range(batch):
                assert got[b] == expect_ntt(vecs[b], inverse, ordering, g), (n, batch, columns, inverse, ordering, g, inplace)


@pytest.mark.parametrize("n,tile_log", [(6, 4), (7, 4), (9, 4), (10, 5), (12, 5), (13, 6)])
def test_ntt_multi_pass_plans(sims, n, tile_log):
    """the 2-, 3- and 4-pass plans (what 2^12..2^29 use with the kernels' 2^11 tile) on small transforms: the harness
    shrinks the tile, the planner and the per-thread bodies are the kernels' own."""
    rng = P.SplitMix64(199 + n)
    for batch, columns in ((1, False), (2, False), (2, True)):
        vecs = [[rng.fr() for _ in range(1 << n)] for _ in range(batch)]
        for inverse, ordering, g, inplace in ((False, 0, 0, True), (True, 0, 0, False), (False, 1, 7, True), (True, 2, 7, True),
                                              (False, 3, 0, False)):
            got = run_ntt(sims, n, batch, inverse, ordering, columns, g, inplace, vecs, tile_log)
            for b in range(batch):
                assert got[b] == expect_ntt(vecs[b], inverse, ordering, g), (n, tile_log, batch, columns, inverse, ordering, g)


@pytest.mark.parametrize("n,log_gpus,a,tile_log", [(8, 1, 4, 11), (10, 2, 5, 11), (10, 3, 5, 4), (12, 1, 6, 5), (12, 3, 6, 5),
                                                    (13, 2, 7, 5)])
def test_ntt_distributed_columns(sims, n, log_gpus, a, tile_log):
    """four-step transform with every rank emulated in turn (column passes with GLOBAL-index twiddles, exchange fused
    into the last pass's stores or done as a transpose, row transforms): the concatenated row blocks are the
    bit-reversed-order (kNR) transform of the whole vector, forward and inverse."""
    rng = P.SplitMix64(299 + n)
    vec = [rng.fr() for _ in range(1 << n)]
    path = os.path.join(sims["dir"], f"ntt_dist_{n}.bin")
    with open(path, "wb") as f:
        f.write(P.fr_bytes(P.fr_to_mont(P.fr_omega(n))))
        f.write(P.fr_bytes(P.fr_to_mont(1)))
        for v in vec:
            f.write(P.fr_bytes(P.fr_to_mont(v)))
    for inverse in (False, True):
        exp = P.apply_ordering(P.ntt(vec, inverse=inverse), "NR", "out")
        for fused in (1, 0):
            r = subprocess.run([sims["ntt_host_sim"], "dist", str(n), str(log_gpus), str(a), str(int(inverse)), str(fused), path,
                                str(tile_log)], capture_output=True, text=True, check=True)
            raw = bytes.fromhex(r.stdout.strip())
            got = [P.fr_from_mont(int.from_bytes(raw[32 * i:32 * i + 32], "little")) for i in range(1 << n)]
            assert got == exp, (n, log_gpus, a, inverse, fused)


# ---------------------------------------------------------------- csrc/glv.cuh: GLV split, scalar multiplication, subgroup checks
@pytest.fixture(scope="module")
def glv_sim(tmp_path_factory):
    out = tmp_path_factory.mktemp("glvsim")
    exe = str(out / "glv_host_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{HOST}", f"-I{CSRC}", "-o", exe, os.path.join(HOST, "glv_host_sim.cpp")],
                   check=True)
    return exe, str(out)


def test_glv_decomposition(glv_sim):
    import vectors_points as V
    exe, d = glv_sim
    ks = V.glv_scalars()
    path = os.path.join(d, "k.bin")
    with open(path, "wb") as f:
        for k in ks:
            f.write(P.fr_bytes(k % P.R_MOD))
    lines = subprocess.run([exe, "decompose", path], capture_output=True, text=True, check=True).stdout.split("\n")
    for k, ln in zip(ks, lines):
        k1, k2 = (int(x, 16) for x in ln.split())
        assert (k1, k2) == P.glv_decompose(k), hex(k)
        assert k1 + k2 * P.GLV_LAMBDA == k % P.R_MOD and k1 < P.GLV_LAMBDA


def test_glv_scalar_mul_bodies(glv_sim):
    import vectors_points as V
    exe, d = glv_sim
    rng = P.SplitMix64(384)
    ks = V.glv_scalars()[:20]
    pts = [P.g1_mul(rng.fr(), P.G1_GEN) for _ in ks]
    pts[3] = None
    path = os.path.join(d, "m.bin")
    with open(path, "wb") as f:
        for k, pt in zip(ks, pts):
            f.write(P.g1_affine_mont_bytes(pt) + P.fr_bytes(k))
    exp = [P.g1_affine_mont_bytes(P.g1_mul(k, pt)).hex() for k, pt in zip(ks, pts)]
    for glv in ("1", "0"):
        got = subprocess.run([exe, "mul", glv, path], capture_output=True, text=True, check=True).stdout.split()
        assert got == exp, glv


def test_subgroup_check_bodies(glv_sim):
    import vectors_points as V
    exe, d = glv_sim
    for group, cases, enc in (("g1", V.g1_membership_cases(), P.g1_affine_mont_bytes),
                              ("g2", V.g2_membership_cases(), P.g2_affine_mont_bytes)):
        assert any(m for _, m in cases) and any(not m for _, m in cases)
        path = os.path.join(d, group + ".bin")
        with open(path, "wb") as f:
            for pt, _ in cases:
                f.write(enc(pt))
        got = subprocess.run([exe, "sub", group, path], capture_output=True, text=True, check=True).stdout.split()
        assert [int(x) for x in got] == [int(m) for _, m in cases], group


def test_binary_gcd_inversion(tmp_path):
    """csrc/inv_bingcd.cuh (Pornin's binary GCD on 64-bit approximations, the inversion behind inv_vartime): y^-1 mod p
    and mod r against Python's pow for edge values (0, 1, m-1, powers of two around the 30/32/64-bit boundaries of the
    approximation, values of every bit length) and 2000 random ones per field; also with one round fewer than shipped,
    which the round-count bound says must still be enough."""
    import random
    exe = str(tmp_path / "inv_host_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{HOST}", f"-I{CSRC}", "-o", exe, os.path.join(HOST, "inv_host_sim.cpp")], check=True)
    rnd = random.Random(20261019)
    for name, mod, nbytes, shipped in (("fq", P.P_MOD, 48, 27), ("fr", P.R_MOD, 32, 18)):
        vals = [0, 1, 2, 3, mod - 1, mod - 2, (mod + 1) // 2, (1 << 30) - 1, 1 << 30, (1 << 31) - 1, 1 << 32, (1 << 64) - 1,
                1 << 64, (1 << 96) + 1, 1 << (mod.bit_length() - 1)]
        vals += [rnd.randrange(1 << k) % mod for k in range(1, mod.bit_length() + 1)]
        vals += [rnd.randrange(mod) for _ in range(2000)]
        path = str(tmp_path / (name + ".bin"))
        with open(path, "wb") as f:
            for v in vals:
                f.write(v.to_bytes(nbytes, "little"))
        for rounds in (0, shipped - 1):
            out = subprocess.run([exe, name, str(rounds), path], capture_output=True, text=True, check=True).stdout.split()
            assert len(out) == len(vals)
            for v, o in zip(vals, out):
                assert int.from_bytes(bytes.fromhex(o), "little") == (pow(v, -1, mod) if v else 0), (name, rounds, hex(v))
