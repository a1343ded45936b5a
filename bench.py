#!/usr/bin/env python3
"""Headline benchmark: G1 MSM 2^24 points/s (+ Fr NTT 2^24 elements/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log-n 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0 (contract in the task statement / DESIGN.md "Measurement").
  * step      = one G1 MSM over 2^24 resident bases with Montgomery scalars (BASELINE.json config 3),
                strong-scaled by contiguous point ranges at N > 1 (one XYZZ partial per GPU, NCCL
                all_gather, combine); the NTT 2^24 is timed right after it and reported under "ntt".
                At N > 1 the line also carries "ntt_fourstep" (Fr NTT 2^26 over the N GPUs, ONE NCCL all_to_all) and
                "plonk_commit_round" (8 MSMs of 2^22 dealt to the ranks), each checked byte for byte against the
                single-GPU result; "result_check" is then sharded MSM == single-GPU MSM (multi_gpu_legs).
  * value     = points/s with scalars already in HBM;  e2e = the reference-facing plugin call itself
                (b381_g1_msm, what msm_cuda_impl is replaced by) with HOST (pinned) scalars, are_scalars_on_device =
                false, synchronous, H2D and the 144-byte result D2H inside the timed region; "e2e_pipelined" = the
                host API's async calls (msm_with_device_bases_async, one stream each) with two commits in flight.
  * roofline  = bucket accumulation (affine pre-reduction levels + k_msm_accumulate) against the IMAD.WIDE issue
                peak MEASURED in this run (the path is integer-pipe bound, not HBM/tensor); "ntt.roofline" is the
                HBM view north_star asks for.
  * cpu_baseline / --impl reference = the CPU port oracle/oracle.c (BLST / midnight-curves are not in
    this image: kind "port") on a bounded sample with every host core.
  * reference_gpu = the reference's OWN sm_100 kernels (oracle/_ref/libref_field.so, libref_msm.so, compiled from
    /root/reference where the sources lie) timed on this GPU beside ours through the same signatures -- the R-GPU
    comparator of SURVEY.md 2.2; outside the product path and outside every timed region of ours.
The oracle is used here ONLY for that CPU leg and for a one-off result check outside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# G1 generator, Montgomery affine (spec constant; same literal as csrc/field_consts.h)
G1_GEN_MONT = [0x5cb38790fd530c16, 0x7817fc679976fff5, 0x154f95c7143ba1c1, 0xf0ae6acdf3d0e747, 0xedce6ecc21dbf440,
               0x120177419e0bfb75, 0xbaac93d50ce72271, 0x8c22631a7918fd8e, 0xdd595f13570725ce, 0x51ac582950405194,
               0x0e1c8c3fad0059c0, 0x0bbc3efc5008a26a]
R_TOP_LIMB = 0x73EDA753299D7D48
# G2 generator, Montgomery affine (x.c0, x.c1, y.c0, y.c1), spec constant
G2_GEN_MONT = [0xf5f28fa202940a10, 0xb3f5fb2687b4961a, 0xa1a893b53e2ae580, 0x9894999d1a3caee9, 0x6f67b7631863366b,
               0x58191924350bcd7, 0xa5a9c0759e23f606, 0xaaa0c59dbccd60c3, 0x3bb17e18e2867806, 0x1b1ab6cc8541b367,
               0xc2b6ed0ef2158547, 0x11922a097360edf3, 0x4c730af860494c4a, 0x597cfa1f5e369c5a, 0xe7e6856caa0a635a,
               0xbbefb5e96e0d495f, 0x7d3a975f0ef25a2, 0x83fd8e7e80dae5, 0xadc0fc92df64b05d, 0x18aa270a2b1461dc,
               0x86adac6a3be4eba0, 0x79495c4ec93da33a, 0xe7175850a43ccaed, 0xb2bc2a163de1bf2]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--cpu-log-n", type=int, default=0, help="log2 size of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference's own kernels (oracle/_ref) timed beside ours")
    ap.add_argument("--dist-ntt-log", type=int, default=26, help="log2 size of the four-step NTT timed at N > 1")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        mhz = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None,
                "sm_max_mhz": float(rows[0][1]) if rows and rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(rows), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ----------------------------------------------------------------------------- CPU (reference arm / baseline)
def cpu_leg(log_n: int, steps: int, warmup: int):
    """oracle.c Pippenger + radix-2 NTT on all host cores; returns dicts for MSM and NTT."""
    from oracle import cref as O
    cores = O.num_threads()
    n = 1 << log_n
    bases = O.gen_series(1, [3, 0, 0, 0], [5, 0, 0, 0], n)
    sc = O.random_fr(0xB12381, n)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.msm(1, sc, bases, scalars_mont=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    msm_s = statistics.mean(times)
    ntt_log = min(log_n + 2, 22)
    a = O.random_fr(7, 1 << ntt_log)
    t0 = time.perf_counter()
    O.ntt(a)
    ntt_s = time.perf_counter() - t0
    return {"cores": cores, "msm_log_n": log_n, "msm_s": msm_s, "msm_pts_per_s": n / msm_s,
            "ntt_log_n": ntt_log, "ntt_s": ntt_s, "ntt_elems_per_s": (1 << ntt_log) / ntt_s}


def workload_config(args, world):
    """The `config` object of BOTH arms (the driver compares them): what is computed, nothing measured."""
    return {"workload": f"G1 MSM n=2^{args.log_n}, bases (1+i)G resident in HBM, uniform Montgomery scalars, signed-digit "
                        f"Pippenger; + Fr NTT 2^{args.log_n} (kNN, in place)",
            "sharding": f"{world} contiguous point ranges, one XYZZ partial per GPU, NCCL all_gather" if world > 1 else "single GPU",
            "l2": "inputs (0.5 GiB scalars + 1.5 GiB bases) exceed the 126 MB L2; no flush needed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers: this arm is the CPU path with EVERY host core, at any N
    os.environ.pop("OMP_NUM_THREADS", None)
    from oracle import cref as O
    O.set_threads(os.cpu_count() or 1)
    log_n = args.cpu_log_n or 18
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_leg(log_n, steps, warmup)
    sample = (f"every step = G1 MSM over 2^{log_n} points of the 2^{args.log_n} workload (oracle/oracle.c, OpenMP, "
              f"{r['cores']} threads); Fr NTT 2^{r['ntt_log_n']} once")
    line = {
        "impl": "reference", "metric": f"g1_msm_2^{args.log_n}_points_per_s", "value": r["msm_pts_per_s"], "unit": "points/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["msm_s"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64 (6x64 Fq / 4x64 Fr Montgomery, 32-bit IMAD limbs)", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "note": "CPU port of the reference's MIDNIGHT_DEVICE=cpu path (Pippenger, OpenMP): BLST/midnight-curves are absent from "
                "this image (SURVEY.md 8c), oracle/oracle.c stands in; points/s is size-normalised, the sample is in cpu_baseline",
        "cpu_baseline": {"value": r["msm_pts_per_s"], "unit": "points/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["msm_pts_per_s"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ntt": {"metric": f"fr_ntt_elems_per_s", "value": r["ntt_elems_per_s"], "unit": "elements/s", "log_n": r["ntt_log_n"]},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def canonical_fr(torch, count, seed):
    """`count` canonical Fr values as [count, 4] int64 device words (top limb below r's top limb)."""
    gen = torch.Generator(device="cuda").manual_seed(seed)
    v = torch.randint(0, 1 << 62, (count, 4), dtype=torch.int64, device="cuda", generator=gen)
    v[:, 3] = torch.randint(0, R_TOP_LIMB, (count,), dtype=torch.int64, device="cuda", generator=gen)
    return v


def multi_gpu_legs(args, torch, dist, np, M, L, D, lib, rank, world, local_rank, n, sc, result, timed, barrier):
    """What only exists at N > 1, on real NCCL (BASELINE.json configs 3 and 5):
    (1) sharded-MSM parity: the all_gather'ed result of the timed step against ONE single-GPU MSM over all points,
        run on rank 0 from the gathered scalars;
    (2) Fr NTT 2^26 four-step over the N GPUs (column stages, ONE all_to_all of row blocks over NVLink, row
        transforms): every rank compares its row block with its own single-GPU kNR transform of the same vector,
        both directions, then the forward transform is timed (max over ranks) with the three phases split by events;
    (3) the commitment MSMs of a k = 22 PLONK prover round: 8 scalar vectors over shared resident bases, dealt
        round-robin to the ranks (no data-path collective; 144-byte results all_gather'ed), checked against rank 0
        computing all eight alone."""
    g = np.array(G1_GEN_MONT, dtype=np.uint64)
    dev_cfg = lib.b381_default_msm_config()
    dev_cfg.are_scalars_on_device = dev_cfg.are_points_on_device = True
    dev_cfg.are_scalars_montgomery_form = dev_cfg.are_points_montgomery_form = True

    # ---- (1)
    shard_check = "skipped (uneven shards)"
    if n % world == 0:
        full_sc = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(full_sc, sc)
        if rank == 0:
            full_bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
            L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(n), L.ptr(full_bases), None), "point_series")
            res = np.zeros(18, dtype=np.uint64)
            L.check(lib.b381_g1_msm(L.ptr(full_sc), L.ptr(full_bases), n, C.byref(dev_cfg), L.ptr(res)), "single-GPU msm")
            shard_check = "ok (sharded result == single-GPU MSM of all points on rank 0, bytes)" \
                if res.tobytes() == result["r"].tobytes() == result["e2e"].tobytes() else "MISMATCH"
            del full_bases
        del full_sc
        torch.cuda.empty_cache()

    # ---- (2)
    log_d = args.dist_ntt_log
    nd, loc = 1 << log_d, (1 << log_d) // world
    ctx = M.GpuNttContext(log_d, device_id=local_rank)
    variants = {}
    oracle_ok = None
    for name, fused in (("fused_p2p", True), ("nccl_all_to_all", False)):
        dn = D.DistributedNtt(log_d, fused=fused)
        ok = True
        for direction in (0, 1):
            x = canonical_fr(torch, nd, 0x26)                     # same vector on every rank
            work = D.column_block_of(x, log_d, rank, world).contiguous()
            rows = dn.forward(work, direction)
            if direction == 0 and fused and rank == 0 and not args.no_cpu_baseline and log_d <= 26:
                # independent check (not our own single-GPU transform): rank 0's row block of the distributed result against
                # the CPU oracle's NTT of the same vector, bit-reversed (kNR order), every byte
                from oracle import cref as O
                nat = O.ntt(x.cpu().numpy().view(np.uint64).reshape(-1, 4))
                nr = O.bit_reverse(nat)
                del nat
                oracle_ok = bool((rows.reshape(-1, 4).cpu().numpy().view(np.uint64) == nr[rank * loc:(rank + 1) * loc]).all())
                del nr
            ctx.ntt_on_device(x.data_ptr(), direction, size=nd, ordering=M.ntt.kNR)
            ok = ok and bool(torch.equal(rows.reshape(-1, 4), x[rank * loc:(rank + 1) * loc]))
            del x, rows
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)

        def ntt_step(marks=None):
            dn.forward(work, 0, marks)
        for _ in range(2):
            ntt_step()
        ms, _, _ = timed(ntt_step, args.steps)
        ms /= args.steps
        evs = []

        def mark(_label):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append(e)
        barrier()
        ntt_step(mark)
        barrier()
        ph = torch.tensor([evs[i].elapsed_time(evs[i + 1]) for i in range(3)], device="cuda")
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
        variants[name] = (ms, [round(float(v), 4) for v in ph], int(flag.item()) == 1, dn.fused)
        sh = dn.shape
        del work
        torch.cuda.empty_cache()
    ms, ph, ok_f, was_fused = variants["fused_p2p"]
    ms_n, ph_n, ok_n, _ = variants["nccl_all_to_all"]
    ntt_dist = {"metric": f"fr_ntt_2^{log_d}_fourstep_elements_per_s", "value": nd / (ms * 1e-3), "unit": "elements/s",
                "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                "shape": f"[2^{sh['a']}][2^{sh['lo']}], {sh['L']} columns then {sh['rows_per_rank']} rows per GPU",
                "exchange": "fused into the last column pass: stores go to the owning GPU's row buffer over NVLink peer memory "
                            "(b381_ntt_dist_columns_p2p), two 4-byte all_reduce barriers" if was_fused else "NCCL all_to_all",
                "phases_ms": {"barrier+columns(+remote stores)": ph[0], "barrier": ph[1], "rows": ph[2]},
                "nccl_all_to_all_variant": {"ms_per_step": ms_n, "value": nd / (ms_n * 1e-3),
                                            "phases_ms": {"columns": ph_n[0], "all_to_all+transpose": ph_n[1], "rows": ph_n[2]}},
                "exchange_bytes_per_gpu": loc * 32 * (world - 1) // world,
                "result_check": "ok (every rank's row block == its single-GPU kNR transform, forward and inverse, both variants)"
                if ok_f and ok_n else f"MISMATCH (fused ok={ok_f}, nccl ok={ok_n})",
                "oracle_check": None if oracle_ok is None else
                ("ok (rank 0's row block == oracle.ntt of the whole vector, bit-reversed, bytes)" if oracle_ok else "MISMATCH")}

    # ---- (3)
    nb, nk = 8, 1 << 22
    bases_k = torch.empty((nk, 12), dtype=torch.int64, device="cuda")
    L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(nk), L.ptr(bases_k), None), "point_series")
    mine = list(range(rank, nb, world))

    def run_commits(ids):
        if not ids:
            return np.zeros((0, 18), dtype=np.uint64)
        scal = torch.stack([canonical_fr(torch, nk, 0xC0 + j) for j in ids]).contiguous()
        cfg = lib.b381_default_msm_config()
        cfg.are_scalars_on_device = cfg.are_points_on_device = True
        cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
        cfg.batch_size, cfg.are_points_shared_in_batch = len(ids), True
        res = np.zeros((len(ids), 18), dtype=np.uint64)

        def go():
            L.check(lib.b381_g1_msm(L.ptr(scal), L.ptr(bases_k), nk, C.byref(cfg), L.ptr(res)), "commit batch")
            return res
        return go

    go = run_commits(mine)
    gathered = {}

    def commit_step():
        part = torch.from_numpy(go().view(np.int64)).cuda() if mine else torch.zeros((0, 18), dtype=torch.int64, device="cuda")
        if nb % world == 0:
            allr = torch.empty((nb, 18), dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(allr, part)
            gathered["r"] = allr
    commits = None
    if nb % world == 0:
        commit_step()
        k = max(1, args.steps // 2)
        cms, _, _ = timed(commit_step, k)
        cms /= k
        check = None
        if rank == 0:
            alone = run_commits(list(range(nb)))().copy()
            got = gathered["r"].cpu().numpy().view(np.uint64).reshape(world, nb // world, 18)
            same = all(got[j % world, j // world].tobytes() == alone[j].tobytes() for j in range(nb))
            check = "ok (== the same 8 commits run on rank 0 alone, bytes)" if same else "MISMATCH"
        commits = {"metric": "plonk_k22_commit_round_points_per_s", "value": nb * nk / (cms * 1e-3), "unit": "points/s",
                   "ms_per_step": cms, "batch": nb, "log_n": 22, "n_gpus": world, "scaling": "strong",
                   "sharding": "commits dealt round-robin, bases replicated, results all_gather'ed", "result_check": check}
    return shard_check, ntt_dist, commits


def reference_gpu_leg(args, torch, np, L, lib, timed, sc, sc_host, bases, n, ours_ntt_ms, ours_vmul_ms, ours_msm):
    """R-GPU comparator (SURVEY.md 2.2 / BASELINE.md 2): the reference's own kernels, compiled for sm_100 from where the
    sources lie (oracle/Makefile -> oracle/_ref/), timed on THIS GPU in THIS run through the reference's flat extern "C"
    API (ntt_kernels.cu:1911-1942, vec_ops.cu:693-838, icicle_curve_api.cu:679-692), beside ours through the same
    signatures.  Test infrastructure: nothing here is on the product path or inside a timed region of ours."""
    out = {"what": "the reference's own sm_100 kernels on this GPU, same buffers and flags as ours"}
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    steps = max(2, min(args.steps, 5))
    f_so, m_so = os.path.join(ref_dir, "libref_field.so"), os.path.join(ref_dir, "libref_msm.so")
    if os.path.exists(f_so):
        ref = C.CDLL(f_so)
        from oracle import pyref as Pr
        root = np.array(Pr.to_limbs(Pr.fr_to_mont(Pr.fr_omega(24)), 4), dtype=np.uint64)   # what its init assumes (:1614-1644)
        if ref.bls12_381_ntt_init_domain_cuda(L.ptr(root), C.byref(L.NTTInitDomainConfig())) == 0:
            nn = min(n, 1 << 24)
            x, y = sc[:nn].clone(), torch.empty_like(sc[:nn])
            cfg = lib.b381_default_ntt_config()
            cfg.are_inputs_on_device = cfg.are_outputs_on_device = True

            def ref_ntt():
                L.check(ref.bls12_381_ntt_cuda(L.ptr(x), nn, 0, C.byref(cfg), L.ptr(y)), "reference ntt")

            def our_ntt():
                L.check(lib.bls12_381_ntt_cuda(L.ptr(x), nn, 0, C.byref(cfg), L.ptr(y)), "our ntt, same signature")
            res = {}
            for name, fn in (("reference", ref_ntt), ("ours", our_ntt)):
                fn()
                fn()
                ms, _, _ = timed(fn, steps)
                res[name] = ms / steps
                res[name + "_sha"] = __import__("hashlib").sha256(y.cpu().numpy().tobytes()).hexdigest()[:16]
            out["ntt"] = {"log_n": nn.bit_length() - 1, "entry": "bls12_381_ntt_cuda, device in / device out (out != in), kNN",
                          "reference_ms": res["reference"], "ours_ms": res["ours"], "ours_in_place_ms": ours_ntt_ms,
                          "speedup": res["reference"] / res["ours"], "same_bytes": res["reference_sha"] == res["ours_sha"]}
            ref.bls12_381_ntt_release_domain_cuda()
            del x, y
        vn = min(n, 1 << 22)
        a, b, o = sc[:vn], sc[vn:2 * vn] if n >= 2 * vn else sc[:vn], torch.empty_like(sc[:vn])
        vcfg = lib.b381_default_vecops_config()
        vcfg.is_a_on_device = vcfg.is_b_on_device = vcfg.is_result_on_device = True
        res = {}
        for name, fn in (("reference", ref.bls12_381_vector_mul), ("ours", lib.bls12_381_vector_mul)):
            def vstep(fn=fn):
                L.check(fn(L.ptr(a), L.ptr(b), C.c_size_t(vn), C.byref(vcfg), L.ptr(o)), "vector_mul")
            vstep()
            vstep()
            ms, _, _ = timed(vstep, steps * 4)
            res[name] = ms / (steps * 4)
            res[name + "_sha"] = __import__("hashlib").sha256(o.cpu().numpy().tobytes()).hexdigest()[:16]
        out["vector_mul"] = {"log_n": vn.bit_length() - 1, "entry": "bls12_381_vector_mul, device operands",
                             "reference_ms": res["reference"], "ours_ms": res["ours"], "speedup": res["reference"] / res["ours"],
                             "same_bytes": res["reference_sha"] == res["ours_sha"], "ours_2^24_ms": ours_vmul_ms}
        torch.cuda.empty_cache()
    else:
        out["ntt"] = out["vector_mul"] = "oracle/_ref/libref_field.so not built (needs /root/reference at build time)"
    if os.path.exists(m_so):
        ref = C.CDLL(m_so)
        from oracle import pyref as Pr
        # the reference's flat API takes INTEGER-form scalars and Montgomery bases (icicle_curve_api.cu:672-676)
        rows = []
        for log_m in (16, 20, 22, 24):
            m = 1 << log_m
            if m > n:
                break
            sci = torch.empty((m, 4), dtype=torch.int64, device="cuda")
            vcfg = lib.b381_default_vecops_config()
            vcfg.is_a_on_device = vcfg.is_result_on_device = True
            L.check(lib.b381_montgomery_convert(L.ptr(sc), C.c_uint64(m), 0, C.byref(vcfg), L.ptr(sci)), "from_mont")
            cfg = lib.b381_default_msm_config()
            cfg.are_scalars_on_device = cfg.are_points_on_device = True
            cfg.are_scalars_montgomery_form, cfg.are_points_montgomery_form = False, True
            r_ref, r_our = np.zeros(18, dtype=np.uint64), np.zeros(18, dtype=np.uint64)
            t0 = time.perf_counter()
            rc = ref.bls12_381_g1_msm_cuda(L.ptr(sci), L.ptr(bases), m, C.byref(cfg), L.ptr(r_ref))
            torch.cuda.synchronize()
            first_s = time.perf_counter() - t0
            if rc != 0:
                rows.append({"log_n": log_m, "reference": f"failed with code {rc}"})
                break
            k = 1 if first_s > 2.0 else 3
            ms_ref, _, _ = timed(lambda: ref.bls12_381_g1_msm_cuda(L.ptr(sci), L.ptr(bases), m, C.byref(cfg), L.ptr(r_ref)), k)
            L.check(lib.bls12_381_g1_msm_cuda(L.ptr(sci), L.ptr(bases), m, C.byref(cfg), L.ptr(r_our)), "our msm, same signature")
            ms_our, _, _ = timed(lambda: lib.bls12_381_g1_msm_cuda(L.ptr(sci), L.ptr(bases), m, C.byref(cfg), L.ptr(r_our)), 3)

            def affine(j):                       # Jacobian Montgomery (X, Y, Z) -> affine integers
                X, Y, Z = (Pr.fq_from_mont(Pr.from_limbs(j[6 * i:6 * i + 6])) for i in range(3))
                if Z == 0:
                    return None
                zi = pow(Z, -1, Pr.P_MOD)
                return (X * zi * zi % Pr.P_MOD, Y * zi * zi * zi % Pr.P_MOD)
            rows.append({"log_n": log_m, "reference_ms": ms_ref / k, "ours_ms": ms_our / 3, "speedup": (ms_ref / k) / (ms_our / 3),
                         "same_point": affine(r_ref) == affine(r_our)})
            del sci
            if first_s > 20.0:
                break
        out["g1_msm"] = {"entry": "bls12_381_g1_msm_cuda, device integer-form scalars, resident Montgomery bases, host result",
                         "sizes": rows, "ours_headline_ms": ours_msm}
    else:
        out["g1_msm"] = "oracle/_ref/libref_msm.so not built (needs /root/reference at build time; ~6 min of cicc)"
    return out


_JSON_FD = None


def _claim_stdout():
    """stdout carries exactly ONE line, the JSON result: everything any library prints there (NCCL's version banner,
    torch warnings) is sent to stderr by pointing fd 1 at fd 2; the result is written to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import midnight_bls12_381_cuda_b200 as M
    from midnight_bls12_381_cuda_b200 import _lib as L
    from midnight_bls12_381_cuda_b200 import dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    M.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version / debug lines go to stderr: stdout is the ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = L.lib()
    os.environ["B381_MSM_TIMING"] = "1"
    n = 1 << args.log_n
    beg, end = D.shard_range(n, rank, world)
    n_loc = end - beg

    # ---- inputs: bases (1+i) G laid down in HBM by the library itself; uniform canonical Montgomery scalars
    g = np.array(G1_GEN_MONT, dtype=np.uint64)
    p0 = g.copy()
    bases = torch.empty((n_loc, 12), dtype=torch.int64, device="cuda")
    if beg:
        # P0 = (1 + beg) G through the reference-named flat entry point, whose result is Jacobian in
        # Montgomery form with Z = R, i.e. (x, y) already is the Montgomery affine point
        s1 = np.array([[beg + 1, 0, 0, 0]], dtype=np.uint64)
        res = np.zeros(18, dtype=np.uint64)
        L.check(lib.bls12_381_g1_msm_cuda(L.ptr(s1), L.ptr(g), 1, C.byref(lib.b381_default_msm_config()), L.ptr(res)), "seed point")
        p0 = res[:12].copy()
    L.check(lib.b381_g1_point_series(L.ptr(p0), L.ptr(g), C.c_uint64(n_loc), L.ptr(bases), None), "point_series")
    gen = torch.Generator(device="cuda").manual_seed(0xB12381 + rank)
    sc = torch.randint(0, 1 << 62, (n_loc, 4), dtype=torch.int64, device="cuda", generator=gen) * 4 + \
        torch.randint(0, 4, (n_loc, 4), dtype=torch.int64, device="cuda", generator=gen)
    sc[:, 3] = torch.randint(0, R_TOP_LIMB, (n_loc,), dtype=torch.int64, device="cuda", generator=gen)
    sc_host = torch.empty((n_loc, 4), dtype=torch.int64).pin_memory()
    sc_host.copy_(sc)
    torch.cuda.synchronize()

    msm = D.ShardedMsm("g1")
    result = {}

    part_buf = torch.empty(D.XYZZ_BYTES["g1"], dtype=torch.uint8, device="cuda")
    gathered = torch.empty((world, D.XYZZ_BYTES["g1"]), dtype=torch.uint8, device="cuda")

    def step_resident():
        part = msm.partial(sc, bases, n_loc, scalars_mont=True, out=part_buf)
        parts = D.gather_partials_into(part, gathered)
        if rank == 0:
            result["r"] = msm.combine(parts)

    # e2e: the plugin call itself.  Host (pinned) scalars, resident bases, are_scalars_on_device = false: the library stages
    # the scalars in (H2D on the caller's stream), runs the MSM and writes the result to host memory -- exactly what ICICLE's
    # dispatcher does through msm_cuda_impl's replacement (icicle_curve_api.cu:243-407).  Synchronous, one call per step.
    host_cfg = lib.b381_default_msm_config()
    host_cfg.are_scalars_on_device, host_cfg.are_points_on_device = False, True
    host_cfg.are_scalars_montgomery_form = host_cfg.are_points_montgomery_form = True
    e2e_part = torch.empty(24, dtype=torch.int64, device="cuda")          # one XYZZ partial (N > 1 only)
    e2e_res = np.zeros(18, dtype=np.uint64)

    def step_e2e():
        if world == 1:
            L.check(lib.b381_g1_msm(L.ptr(sc_host), L.ptr(bases), n_loc, C.byref(host_cfg), L.ptr(e2e_res)), "plugin msm")
            result["e2e"] = e2e_res
        else:
            L.check(lib.b381_g1_msm_partial(L.ptr(sc_host), L.ptr(bases), n_loc, C.byref(host_cfg), L.ptr(e2e_part)), "plugin partial")
            parts = D.gather_partials_into(e2e_part.view(torch.uint8), gathered)
            if rank == 0:
                result["e2e"] = msm.combine(parts)               # D2H of the 144-byte result

    # e2e_pipelined (N = 1): the host API's async calls, each on its own stream with its own H2D
    # (core/msm.rs:715-798, :1314-1418): two commits in flight, so the copy of one rides under the other's kernels.
    def run_pipelined(k):
        ctx = M.GpuMsmContext(device_id=local_rank)
        from midnight_bls12_381_cuda_b200.stream import DeviceVec
        dev_bases = M.msm.PrecomputedBases(DeviceVec.borrow(bases.data_ptr(), n_loc, 96), n_loc)
        sc_np = sc_host.numpy().view(np.uint64)
        pending, out = [], None
        for _ in range(k):
            pending.append(ctx.msm_with_device_bases_async(sc_np, dev_bases))
            if len(pending) == 2:
                out = pending.pop(0).wait()
        while pending:
            out = pending.pop(0).wait()
        return out

    # value_pipelined: the same resident MSM with TWO steps in flight on two streams (async C-ABI calls), the way a prover
    # issues its commitments:
    # the latency-bound tail of one step (bucket reduction, window combine: ~3 ms whatever n is) runs under the affine
    # levels of the next.  One host synchronisation at the end; every result is compared with the serial step's.
    pipe_streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    pipe_bufs = [(torch.empty(D.XYZZ_BYTES["g1"], dtype=torch.uint8, device="cuda"),
                  torch.empty((world, D.XYZZ_BYTES["g1"]), dtype=torch.uint8, device="cuda"),
                  torch.zeros(18, dtype=torch.int64, device="cuda")) for _ in range(2)]
    res_cfg = lib.b381_default_msm_config()
    res_cfg.are_scalars_on_device = res_cfg.are_points_on_device = res_cfg.are_results_on_device = True
    res_cfg.are_scalars_montgomery_form = res_cfg.are_points_montgomery_form = True
    res_cfg.is_async = True

    def run_resident_pipelined(k):
        done = [None, None]
        allp = torch.empty((k, D.XYZZ_BYTES["g1"]), dtype=torch.uint8, device="cuda") if world > 1 else None
        for i in range(k):
            s = pipe_streams[i & 1]                # step i follows step i - 2 on its stream, so their buffers can be shared
            res = pipe_bufs[i & 1][2]
            if done[i & 1] is not None:
                done[i & 1].synchronize()          # at most two steps in flight (what a handle's wait() does for a prover)
            with torch.cuda.stream(s):
                if world == 1:
                    res_cfg.stream = C.c_void_p(s.cuda_stream)
                    L.check(lib.b381_g1_msm(L.ptr(sc), L.ptr(bases), n_loc, C.byref(res_cfg), L.ptr(res)), "async msm")
                else:
                    msm.partial(sc, bases, n_loc, scalars_mont=True, stream=s.cuda_stream, out=allp[i])
                done[i & 1] = torch.cuda.Event()
                done[i & 1].record(s)
        torch.cuda.synchronize()
        if world == 1:
            return [b[2].cpu().numpy().view(np.uint64).copy() for b in pipe_bufs[:min(k, 2)]]
        # N > 1: ONE all_gather for the k partials of every rank, then k stream-ordered combines.  A collective per step
        # inside the pipeline couples the ranks step by step (every all_gather waits for the slowest rank while the next
        # one queues behind it on NCCL's stream): measured 15.2 ms per step on 8 GPUs against 12.6 serial and 11.0 this way
        # (tools/gpu_pipe_nccl.py, profiles/r02h_value_pipelined.txt)
        allg = torch.empty((world, k, D.XYZZ_BYTES["g1"]), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allg.view(-1), allp.view(-1))
        allr = torch.zeros((k, 18), dtype=torch.int64, device="cuda")
        for i in range(k):
            msm.combine_async(allg[:, i].contiguous(), allr[i])
        torch.cuda.synchronize()
        return [r.numpy().view(np.uint64).copy() for r in allr.cpu()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, phase_sink=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            fn()
            if phase_sink is not None:
                buf = (C.c_float * 12)()
                cnt = lib.b381_msm_last_timings(buf, 12)
                row = [buf[i] for i in range(cnt)]
                f_ms, b_ms = C.c_float(), C.c_float()
                if lib.b381_msm_last_level0_ms(C.byref(f_ms), C.byref(b_ms)) == 1:
                    row += [f_ms.value, b_ms.value]
                phase_sink.append(row)
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    phases = []
    ms_total, t0, t1 = timed(step_resident, args.steps, phases)
    msm_info = (C.c_int * 4)()
    lib.b381_msm_last_info(msm_info, 4)          # shape of the G1 MSM just timed (before any other MSM runs)
    clocks = sampler.window(t0, t1) if sampler else None
    if sampler:
        sampler.stop()       # polled during the headline region only: the host-latency-sensitive legs below run undisturbed
    # phase events make every call wait for its own completion (PhaseTimer::finish): off for the end-to-end legs, or
    # the "async" handles would serialise
    os.environ["B381_MSM_TIMING"] = "0"
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    # N = 1 only: inside this script the same leg at N > 1 measured erratic (2 GPUs: 49.9 / 163 ms per step against 39.1
    # serial; 8 GPUs: 20.2 against 12.8) although the stand-alone driver of exactly this scheme, tools/gpu_pipe_nccl.py,
    # measures 36.6 ms at 2 x 2^23 and 11.0 ms at 8 x 2^21 (serial 38.9 / 12.6): not understood, recorded in
    # profiles/r02h_value_pipelined.txt, and not reported here.
    ms_rpipe, rpipe_res = None, []
    if world == 1:
        run_resident_pipelined(args.steps)      # untimed: the pool grows to two working sets once
        box2 = {}
        ms_rpipe, _, _ = timed(lambda: box2.__setitem__("r", run_resident_pipelined(args.steps)), 1)
        rpipe_res = box2["r"]
    ms_pipe, pipe_xy = None, None
    if world == 1:
        run_pipelined(max(4, args.steps))   # the pool has to grow to two working sets once; not part of the measurement
        box = {}
        ms_pipe, _, _ = timed(lambda: box.__setitem__("xy", run_pipelined(args.steps)), 1)
        pipe_xy = box["xy"]

    # ---- NTT 2^24: every rank transforms its own resident vector (replicas; the four-step transform is in multi_gpu_legs)
    ntt_ctx = M.GpuNttContext(args.log_n, device_id=local_rank)
    vec = sc.clone()                                   # canonical Montgomery words
    vec_host = sc_host
    ntt_out_host = torch.empty_like(sc_host).pin_memory() if n_loc == n else None
    ntt_size = n_loc if n_loc & (n_loc - 1) == 0 else n

    def ntt_step():
        ntt_ctx.ntt_on_device(vec.data_ptr(), 0, size=ntt_size)

    ntt_ok = n_loc & (n_loc - 1) == 0
    ntt = None
    if ntt_ok:
        # one-off check outside the timed region: every byte of the forward kNN transform against the CPU oracle
        ntt_check = "skipped"
        if rank == 0 and not args.no_cpu_baseline and ntt_size <= (1 << 26):
            from oracle import cref as O
            ntt_step()
            got = vec.cpu().numpy().view(np.uint64).reshape(-1, 4)
            exp = O.ntt(sc_host.numpy().view(np.uint64).reshape(-1, 4)[:ntt_size])
            ntt_check = "ok" if (got[:ntt_size] == exp).all() else "MISMATCH"
            del got, exp
            vec.copy_(sc)
        for _ in range(args.warmup):
            ntt_step()
        ntt_ms, _, _ = timed(ntt_step, args.steps)
        ntt_ms /= args.steps
        ntt_e2e = None
        if ntt_out_host is not None:
            # the plugin call with HOST buffers: b381_ntt stages in, transforms, copies out (icicle_field_api.cu:97-131)
            ncfg = lib.b381_default_ntt_config()
            ncfg.are_inputs_on_device = ncfg.are_outputs_on_device = False

            def ntt_e2e_step():
                L.check(lib.b381_ntt(L.ptr(vec_host), ntt_size, 0, C.byref(ncfg), L.ptr(ntt_out_host)), "plugin ntt")
            ntt_e2e_step()
            k = max(1, args.steps // 2)
            ms, _, _ = timed(ntt_e2e_step, k)
            ntt_e2e = ms / k
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ninfo = (C.c_int * 4)()
        passes = lib.b381_ntt_last_info(ninfo, 4) and ninfo[0]
        alg_bytes = 128.0 * n_loc          # 2^13..2^24 -> 128 B/element (two ideal passes)
        ntt = {"metric": f"fr_ntt_2^{n_loc.bit_length() - 1}_elements_per_s", "value": n_loc * world / (ntt_ms * 1e-3),
               "unit": "elements/s", "ms_per_step": ntt_ms, "scaling": "weak (one resident transform per GPU)" if world > 1 else "single",
               "e2e": None if ntt_e2e is None else {"value": n_loc / (ntt_e2e * 1e-3), "unit": "elements/s", "ms_per_step": ntt_e2e,
                                                     "h2d_bytes_per_step": n_loc * 32, "d2h_bytes_per_step": n_loc * 32,
                                                     "note": "b381_ntt with pinned HOST input and output (are_inputs_on_device = "
                                                             "are_outputs_on_device = false); PCIe-bound: 2 x 0.5 GiB per transform"},
               "roofline": {"bound": "hbm", "achieved": alg_bytes / (ntt_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": alg_bytes / (ntt_ms * 1e-3) / 1e9 / hbm_peak,
                            "traffic": 3.659e9 * (n_loc / float(1 << 24)) if passes == 3 else None,
                            "traffic_source": "dram__bytes_read+write of the three k_ntt launches of one 2^24 transform (1.58 + 1.02 + 1.06 GB), "
                                              "ncu --set full (profiles/r02h_ntt_key_metrics.txt), scaled by n; achieved/traffic are per transform",
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                            "kernel": f"k_ntt x{passes}", "note": "compute (IMAD) bound: see imad_frac / fr_mul_floor_frac"},
               "result_check": ntt_check, "gpu_launches": passes}

    # ---- vecops 2^24 (96 B/element a o b, 64 B scalar o vec), resident operands, HBM roofline
    vec_ops = None
    if world == 1 and ntt is not None:
        va, vb, vo = sc, vec, torch.empty_like(sc)
        vcfg = lib.b381_default_vecops_config()
        vcfg.is_a_on_device = vcfg.is_b_on_device = vcfg.is_result_on_device = True
        one_elem = sc[:1].clone()
        vec_ops = {}
        for name, fn, a0, bytes_per in (("vector_mul", lib.b381_vector_mul, va, 96.0), ("vector_add", lib.b381_vector_add, va, 96.0),
                                        ("scalar_mul_vec", lib.b381_scalar_mul_vec, one_elem, 64.0)):
            def vstep(fn=fn, a0=a0):
                L.check(fn(L.ptr(a0), L.ptr(vb), C.c_uint64(n_loc), C.byref(vcfg), L.ptr(vo)), name)
            for _ in range(3):
                vstep()
            ms, _, _ = timed(vstep, args.steps)
            ms /= args.steps
            gbs = bytes_per * n_loc / (ms * 1e-3) / 1e9
            vec_ops[name] = {"ms": ms, "elements_per_s": n_loc / (ms * 1e-3),
                             "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                          "bytes_per_element": bytes_per, "kernel": "k_vecop"}}
        del vo

    # ---- roofline denominators measured on this box, now (rank 0)
    # ---- G2 MSM 2^20 (BASELINE.json config 4), single GPU only: bases (1+i) G2 resident, first 2^20 scalars
    g2 = None
    if world == 1 and args.log_n >= 20:
        n2 = 1 << 20
        g2g = np.array(G2_GEN_MONT, dtype=np.uint64)
        bases2 = torch.empty((n2, 24), dtype=torch.int64, device="cuda")
        L.check(lib.b381_g2_point_series(L.ptr(g2g), L.ptr(g2g), C.c_uint64(n2), L.ptr(bases2), None), "g2 point_series")
        cfg2 = lib.b381_default_msm_config()
        cfg2.are_scalars_on_device = cfg2.are_points_on_device = True
        cfg2.are_scalars_montgomery_form = cfg2.are_points_montgomery_form = True
        res2 = np.zeros(36, dtype=np.uint64)

        def g2_step():
            L.check(lib.b381_g2_msm(L.ptr(sc), L.ptr(bases2), n2, C.byref(cfg2), L.ptr(res2)), "g2 msm")
        os.environ["B381_MSM_TIMING"] = "1"
        for _ in range(2):
            g2_step()
        g2_phases = []
        g2_ms, _, _ = timed(g2_step, args.steps, g2_phases)
        g2_ms /= args.steps
        g2_info = (C.c_int * 4)()
        lib.b381_msm_last_info(g2_info, 4)
        g2_info = list(g2_info)
        g2 = {"metric": "g2_msm_2^20_points_per_s", "value": n2 / (g2_ms * 1e-3), "unit": "points/s", "ms_per_step": g2_ms,
              "phases_ms": [round(statistics.mean(c), 4) for c in zip(*g2_phases)][:8], "result": res2}
        del bases2

    # ---- N > 1 only: the exchanges of BASELINE.json configs 3 and 5 on real NCCL
    shard_check, ntt_dist, commits = None, None, None
    if world > 1:
        shard_check, ntt_dist, commits = multi_gpu_legs(args, torch, dist, np, M, L, D, lib, rank, world, local_rank, n, sc, result,
                                                        timed, barrier)

    out = None
    if rank == 0:
        v, ms = C.c_double(), C.c_float()
        L.check(lib.b381_bench_imad_peak(4000, C.byref(v), C.byref(ms)), "imad probe")
        imad_peak = v.value
        L.check(lib.b381_bench_field_mul(0, 1000, C.byref(v), C.byref(ms)), "fq probe")
        fq_rate = v.value
        L.check(lib.b381_bench_field_mul(1, 1000, C.byref(v), C.byref(ms)), "fr probe")
        fr_rate = v.value
        ms_step = ms_total / args.steps
        ph = [statistics.mean(c) for c in zip(*phases)] if phases else []
        names = ["sort (histogram+scan+scatter)", "(unused 1)", "(unused 2)", "prereduce", "tasks+accumulate", "finalize", "bucket_reduce", "combine"]
        c_win, W, levels, own_launches = msm_info[0], msm_info[1], msm_info[2], msm_info[3]
        # Dominant stage = bucket accumulation: `levels` affine pre-reduction levels (k_msm_pair_fwd / k_msm_invert_totals /
        # k_msm_pair_bwd, csrc/msm_batch.cuh) + k_msm_accumulate on what is left.  Algorithmic work per (point, window)
        # insertion as SURVEY.md 8d defines it: one XYZZ mixed addition = 10 Fq products x 300 32x32->64 multiply-adds.
        # The affine levels EXECUTE 6 products per insertion instead; "executed_frac" is the pipe-level figure.
        acc_ms = (ph[3] + ph[4]) if len(ph) > 4 else None
        roof = None
        if acc_ms and W:
            ins = float(n_loc) * W
            ach = ins * 3000.0 / (acc_ms * 1e-3)
            frac_removed = 1.0 - 0.5 ** levels
            executed = ins * (frac_removed * 6 + (1.0 - frac_removed) * 10) * 300.0 / (acc_ms * 1e-3)
            roof = {"bound": "imad (integer pipe; neither HBM nor tensor)",
                    "kernel": f"bucket accumulation: {levels} x (k_msm_pair_fwd, k_msm_invert_totals, k_msm_pair_bwd) + k_msm_accumulate<fq_t>",
                    "achieved": ach / 1e9, "peak": imad_peak / 1e9, "unit": "GMAD/s", "frac": ach / imad_peak,
                    "executed_frac": executed / imad_peak,
                    "algorithmic_unit": "3000 MAD per (point, window) insertion (SURVEY.md 8d); executed: 1800 in the affine levels",
                    "traffic": None, "peak_source": "b381_bench_imad_peak, this run", "kernel_ms": acc_ms,
                    "share_of_step": acc_ms / ms_step, "window_c": c_win, "windows": W, "affine_levels": levels}
            if len(ph) >= 10 and levels:
                # the single dominant kernel: level-0 backward pass, 5 Fq products (1500 MADs) per pair sum, n*W/2 pairs
                pairs = ins / 2.0
                roof["dominant_kernel"] = {
                    "name": "k_msm_pair_bwd<fq_t, 32, level 0>", "kernel_ms": ph[9], "share_of_step": ph[9] / ms_step,
                    "achieved": pairs * 1500.0 / (ph[9] * 1e-3) / 1e9, "unit": "GMAD/s", "frac": pairs * 1500.0 / (ph[9] * 1e-3) / imad_peak,
                    "note": "carry-chain IMAD.WIDE.X issues at ~0.57 of the plain IMAD.WIDE peak used as denominator (profiles/r01_imad_variants.txt)",
                    "traffic": 61.23e9 * (n_loc / float(1 << 24)),
                    "traffic_source": "dram__bytes_read+write of this kernel at n = 2^24 (45.63 + 15.60 GB), ncu --set full (profiles/r02h_msm_key_metrics.txt), scaled by n; algorithmic: 2 x 96 B gathered + 96 B written per pair sum = 38.7 GB",
                    "level0_fwd_ms": ph[8]}
        if ntt:
            # n/2 * log2(n) butterflies, one Fr Montgomery product (2*8^2+8 = 136 multiply-adds) each
            ln = n_loc.bit_length() - 1
            ntt["roofline"]["imad_frac"] = (0.5 * n_loc * ln * 136) / (ntt["ms_per_step"] * 1e-3) / imad_peak
            # floor = n/2 * log2(n) Fr products at the fr_mul rate measured by the probe in this run
            ntt["roofline"]["fr_mul_floor_ms"] = 0.5 * n_loc * ln / fr_rate * 1e3
            ntt["roofline"]["fr_mul_floor_frac"] = ntt["roofline"]["fr_mul_floor_ms"] / ntt["ms_per_step"]
        # one-off correctness check outside the timed region (oracle as checker): sum s_i (beg+1+i) G
        rpipe_check = "ok (every in-flight result == the serial step's, bytes)" \
            if all(r.tobytes() == result["r"].tobytes() for r in rpipe_res) else "MISMATCH"
        check = "skipped"
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import cref as O
                from oracle import pyref as Pr
                s_np = sc_host.numpy().view(np.uint64)
                kk = np.zeros((n, 4), dtype=np.uint64)
                kk[:, 0] = np.arange(1, n + 1, dtype=np.uint64)
                dl = Pr.from_limbs(O.fr_dot(s_np, kk, s_mont=True))
                exp = Pr.g1_result_std_bytes(Pr.g1_mul(dl, Pr.G1_GEN))
                pipe_ok = pipe_xy is None or (pipe_xy[0].to_bytes(48, "little") + pipe_xy[1].to_bytes(48, "little")) == exp[:96]
                check = "ok" if result["r"].tobytes() == exp and result["e2e"].tobytes() == exp and pipe_ok else "MISMATCH"
                if g2 is not None:
                    kk2 = np.zeros((1 << 20, 4), dtype=np.uint64)
                    kk2[:, 0] = np.arange(1, (1 << 20) + 1, dtype=np.uint64)
                    dl2 = Pr.from_limbs(O.fr_dot(s_np[:1 << 20], kk2, s_mont=True))
                    exp2 = Pr.g2_result_std_bytes(Pr.g2_mul(dl2, Pr.G2_GEN))
                    g2["result_check"] = "ok" if g2["result"].tobytes() == exp2 else "MISMATCH"
            except Exception as e:  # noqa: BLE001
                check = f"error: {e}"
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cl = args.cpu_log_n or 20
            r = cpu_leg(cl, 3, 1)
            cpu = {"value": r["msm_pts_per_s"], "unit": "points/s", "cores": r["cores"], "kind": "port",
                   "sample": f"G1 MSM 2^{cl} of the 2^{args.log_n} workload x3 after one warm-up (oracle/oracle.c, OpenMP); NTT 2^{r['ntt_log_n']}: "
                             f"{r['ntt_elems_per_s']:.3e} elements/s"}
        ref_gpu = None
        if world == 1 and not args.no_reference_gpu:
            try:
                ref_gpu = reference_gpu_leg(args, torch, np, L, lib, timed, sc, sc_host, bases, n,
                                            ntt["ms_per_step"] if ntt else None,
                                            vec_ops["vector_mul"]["ms"] if vec_ops else None, ms_step)
            except Exception as e:  # noqa: BLE001
                ref_gpu = {"error": repr(e)}
        out = {
            "metric": f"g1_msm_2^{args.log_n}_points_per_s", "value": n / (ms_step * 1e-3), "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64 (6x64 Fq / 4x64 Fr Montgomery, 32-bit IMAD limbs)",
            "data": "synthetic",
            "config": workload_config(args, world),
            "msm_shape": {"window_c": c_win, "windows": W, "affine_levels": levels},
            "e2e": {"value": n / (ms_e2e / args.steps * 1e-3), "unit": "points/s", "h2d_bytes_per_step": n_loc * 32,
                    "d2h_bytes_per_step": 144, "ms_per_step": ms_e2e / args.steps,
                    "note": "the plugin call: b381_g1_msm" + ("" if world == 1 else "_partial + all_gather + combine") +
                            " with pinned HOST scalars (are_scalars_on_device = false), resident bases, synchronous, "
                            "result to host"},
            "e2e_pipelined": None if ms_pipe is None else {
                "value": n / (ms_pipe / args.steps * 1e-3), "unit": "points/s", "ms_per_step": ms_pipe / args.steps,
                "h2d_bytes_per_step": n_loc * 32, "d2h_bytes_per_step": 144,
                "note": "GpuMsmContext.msm_with_device_bases_async (own stream per call, host scalars staged by the plugin call), "
                        "two commits in flight"},
            "value_pipelined": None if ms_rpipe is None else {
                "value": n / (ms_rpipe / args.steps * 1e-3), "unit": "points/s", "ms_per_step": ms_rpipe / args.steps,
                "result_check": rpipe_check,
                "note": "same resident workload, the K steps issued as async C-ABI calls on two alternating streams (two MSMs in "
                        "flight" + ("" if world == 1 else "; the K partials of every rank are exchanged by ONE all_gather at the end, then K combines") +
                        "), step i + 2 issued when step i has completed (what a handle's wait() does): the latency-bound tail "
                        "of one step runs under the affine levels of the next.  Reported beside `value`, which times the steps "
                        "one after the other"},
            "gpu_launches": (own_launches + 1) * args.steps,
            "gpu_launches_note": "own kernels per MSM step as counted by the library (b381_msm_last_info): histogram, scan "
                                 "(tile / totals / add), scatter, 4 + scan per affine level, task_count/build_tasks, task order "
                                 "(hist / scan / scatter), accumulate, finalize, segment, tree levels, combine, + encode; "
                                 "no library kernel is on the MSM path",
            "phases_ms": dict(zip(names, [round(x, 4) for x in ph[:8]])),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "ntt": ntt, "vecops": vec_ops, "reference_gpu": ref_gpu,
            "g2": None if g2 is None else dict(
                {k: v for k, v in g2.items() if k != "result"},
                msm_shape={"window_c": g2_info[0], "windows": g2_info[1], "affine_levels": g2_info[2]},
                roofline=None if len(g2["phases_ms"]) < 5 else {
                    "bound": "imad (integer pipe)", "kernel": "bucket accumulation: affine levels + k_msm_accumulate<fq2_t>",
                    "algorithmic_unit": "9000 MAD per (point, window) insertion: one XYZZ mixed addition = 10 Fq2 products = 30 Fq "
                                        "products (SURVEY.md 8d's G1 unit x 3); executed in the affine levels: 6 Fq2 products = 5400",
                    "kernel_ms": g2["phases_ms"][3] + g2["phases_ms"][4],
                    "achieved": (1 << 20) * g2_info[1] * 9000.0 / ((g2["phases_ms"][3] + g2["phases_ms"][4]) * 1e-3) / 1e9,
                    "peak": imad_peak / 1e9, "unit": "GMAD/s",
                    "frac": (1 << 20) * g2_info[1] * 9000.0 / ((g2["phases_ms"][3] + g2["phases_ms"][4]) * 1e-3) / imad_peak,
                    "executed_frac": (1 << 20) * g2_info[1] * 5400.0 / ((g2["phases_ms"][3] + g2["phases_ms"][4]) * 1e-3) / imad_peak,
                    "tail_ms": sum(g2["phases_ms"][5:8]),
                    "note": "finalize + bucket reduction + window combine (tail_ms) are latency-bound chains of Fq2 products; "
                            "ncu: profiles/r02e_g2_key_metrics.txt"}),
            "probes": {"imad_wide_mad_per_s": imad_peak, "fq_mul_per_s": fq_rate, "fr_mul_per_s": fr_rate},
            "result_check": check if world == 1 else shard_check,
            "ntt_fourstep": ntt_dist, "plonk_commit_round": commits,
        }
    if sampler:
        sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


if __name__ == "__main__":
    main()
