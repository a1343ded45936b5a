// CPU-only driver for the per-thread MSM bodies in csrc/msm_core.cuh (TEST INFRASTRUCTURE).
// Runs exactly the kernel pipeline of csrc/msm_impl.cuh with a serial loop per "kernel" and
// a serial exclusive scan in place of the device scan.  Usage:
//   msm_host_sim <g1|g2> <n> <c> <K> <L> <scalars_mont 0|1> <infile> [factor] [levels] [batch] [shared 0|1] [chunk_log] [piece_chunks]
// chunk_log < 31 (with levels >= 1) groups the entries chunk-major exactly as msm_impl.cuh does for large inputs.
// prints one result per MSM of the batch (hex, std form).  batch > 1 folds the MSMs into one pipeline run exactly as
// msm_impl.cuh does (scalars [batch][n]; points shared or [batch][n]).
// levels > 0 runs that many affine pre-reduction levels (csrc/msm_batch.cuh) before the tasks, 3 output
// slots per simulated thread; forward, batched inversion of the thread totals and backward run as the three
// per-level kernels do.
// infile = n*32 B scalars followed by n*(96|192) B Montgomery affine points.
// With factor > 1 the bases are first expanded exactly like k_precompute_bases does.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#define B381_HOST_TEST 1
#include "msm_core.cuh"
using namespace b381;

template <class F>
int run(uint32_t n, uint32_t c, uint32_t K, uint32_t L, bool mont, FILE* f, uint32_t factor, uint32_t levels,
        uint32_t batch, bool shared, uint32_t chunk_log, uint32_t piece_chunks) {
  const uint32_t nsc = n * batch, np = shared ? n : n * batch;
  std::vector<fr_t> sc(nsc);
  std::vector<affine_t<F>> pts(np);
  if (fread(sc.data(), sizeof(fr_t), nsc, f) != nsc) return 2;
  if (fread(pts.data(), sizeof(affine_t<F>), np, f) != np) return 2;
  msm_shape sh = make_msm_shape(n, c, 255, factor, batch, shared);
  if (levels >= 1) msm_shape_set_chunks(sh, chunk_log);
  const uint32_t nruns = msm_runs(sh);
  if (factor > 1) {
    std::vector<affine_t<F>> ex((size_t)np * factor);
    for (uint32_t i = 0; i < np; i++) {
      ex[(size_t)i * factor] = pts[i];
      xyzz_t<F> acc = to_xyzz(pts[i]);
      for (uint32_t k = 1; k < factor; k++) {
        for (uint32_t b = 0; b < c * sh.Wf; b++) acc = xyzz_dbl(acc);
        ex[(size_t)i * factor + k] = xyzz_to_affine(acc);
      }
    }
    pts.swap(ex);
  }
  size_t total = (size_t)nsc * sh.W;
  // counting sort exactly as msm_sort.cu runs it: histogram, exclusive scan, scatter through per-slot cursors.  The
  // "threads" of the scatter pass run in a scrambled order (odd n: descending) because the device gives no ordering
  // guarantee inside a bucket and nothing downstream may depend on one.
  std::vector<uint32_t> hist(nruns + 1, 0), offsets(nruns + 1, 0xdeadbeef), sv(total, 0xdeadbeef);
  for (uint32_t i = 0; i < nsc; i++) msm_hist_body(i, sc.data(), mont, sh, hist.data());
  uint32_t run = 0;
  for (uint32_t b = 0; b <= nruns; b++) { offsets[b] = run; run += hist[b]; }
  if (offsets[nruns] != total) return 4;
  std::vector<uint32_t> cursor(offsets);
  for (uint32_t i = 0; i < nsc; i++) msm_scatter_body((n & 1) ? nsc - 1 - i : i, sc.data(), mont, sh, cursor.data(), sv.data());
  for (uint32_t b = 0; b < nruns; b++) if (cursor[b] != offsets[b + 1]) return 5;
  // affine pre-reduction levels
  const uint32_t* cur_vals = sv.data();
  level_pts<F> cur = level_from_bases<F>(pts.data());
  std::vector<F> lvx, lvy;
  for (uint32_t l = 0; l < levels; l++) {
    constexpr int PB = 3;
    const bool chunked = l == 0 && sh.nchunks > 1;
    const uint32_t nb_l = chunked ? nruns : sh.nbuckets;
    std::vector<uint32_t> half(nb_l + 1), next_off(nb_l + 1), half_bm, dst_base;
    if (chunked) {
      half_bm.assign(nb_l + 1, 0xdeadbeef);
      dst_base.assign(nb_l + 1, 0);
      for (uint32_t r = 0; r <= nb_l; r++) msm_half_counts_runs_body(r, offsets.data(), sh.nbuckets, msm_set_slots(sh), sh.nchunks, sh.Bs, half.data(), half_bm.data());
      uint32_t acc2 = 0;
      for (uint32_t r = 0; r <= nb_l; r++) { dst_base[r] = acc2; acc2 += half_bm[r]; }
    } else {
      for (uint32_t b = 0; b <= nb_l; b++) msm_half_counts_body(b, offsets.data(), sh.nbuckets, sh.Bs, half.data());
    }
    uint32_t run_sum = 0;
    for (uint32_t b = 0; b <= nb_l; b++) { next_off[b] = run_sum; run_sum += half[b]; }
    uint32_t n_out = next_off[nb_l];
    std::vector<F> outx((size_t)n_out + PB), outy((size_t)n_out + PB);
    // the three kernels of a level, thread by thread, with the same slot-major global layout
    uint32_t NT = (n_out + PB - 1) / PB + 2;
    std::vector<uint32_t> srcg((size_t)NT * PB), dstg((size_t)NT * PB, 0xdeadbeef);
    std::vector<F> preg((size_t)NT * PB), tot(NT);
    if (chunked && piece_chunks) {
      // streamed level 0 as msm_impl.cuh runs it under the copy of host scalars: per piece of `piece_chunks` chunks the
      // half counts of its runs (+ sentinel), a scan continued through a carried slot total, and the forward pass of the
      // threads the piece completes -- each seeing only the offsets [0, r1] that exist by then; the bucket-major
      // destinations follow in a second walk without loads
      std::vector<uint32_t> h2(nb_l + 1, 0xdeadbeef), hb2(nb_l + 1, 0xdeadbeef), oo(nb_l + 1, 0xdeadbeef);
      std::vector<int> ran(NT, 0);
      uint32_t carry = 0;
      for (uint32_t c0 = 0; c0 < sh.nchunks; c0 += piece_chunks) {
        const uint32_t r0 = c0 * sh.nbuckets;
        const bool last = c0 + piece_chunks >= sh.nchunks;
        const uint32_t r1 = last ? nruns : (c0 + piece_chunks) * sh.nbuckets;
        for (uint32_t r = r0; r <= r1; r++)
          msm_half_counts_runs_body(r, offsets.data(), sh.nbuckets, msm_set_slots(sh), sh.nchunks, sh.Bs, h2.data(), hb2.data(), r1);
        const uint32_t lo = carry;
        for (uint32_t r = r0; r <= r1; r++) { oo[r] = carry; carry += h2[r]; }
        const uint32_t hi = carry;
        if (oo[r1] != hi) return 6;
        for (uint32_t t = 0; t < NT; t++) {
          if (!pair_piece_owns(t, PB, lo, hi, last)) continue;
          if (ran[t]++) return 7;
          pair_walk<PB>(t * PB, hi, offsets.data(), oo.data(), r1, srcg.data() + t, NT);
          tot[t] = pair_phase1<F, PB, true>(srcg.data() + t, NT, cur_vals, cur, preg.data() + t, NT);
        }
      }
      for (uint32_t r = 0; r <= nb_l; r++) if (oo[r] != next_off[r] || h2[r] != half[r] || hb2[r] != half_bm[r]) return 8;
      for (uint32_t t = 0; t < NT; t++) {
        if (((size_t)t * PB < n_out) != (ran[t] == 1)) return 9;
        if ((size_t)t * PB < n_out)
          pair_walk<PB>(t * PB, n_out, offsets.data(), next_off.data(), nb_l, nullptr, NT, dst_base.data(), msm_set_slots(sh), sh.nchunks, dstg.data() + t);
      }
    } else
    for (uint32_t t = 0; t < NT; t++) {
      if (chunked) pair_walk<PB>(t * PB, n_out, offsets.data(), next_off.data(), nb_l, srcg.data() + t, NT, dst_base.data(), msm_set_slots(sh), sh.nchunks, dstg.data() + t);
      else pair_walk<PB>(t * PB, n_out, offsets.data(), next_off.data(), nb_l, srcg.data() + t, NT);
      tot[t] = l == 0 ? pair_phase1<F, PB, true>(srcg.data() + t, NT, cur_vals, cur, preg.data() + t, NT)
                      : pair_phase1<F, PB, false>(srcg.data() + t, NT, nullptr, cur, preg.data() + t, NT);
    }
    constexpr int M = 8;
    uint32_t mm = 1 + l % M;
    uint32_t T2 = (NT + mm - 1) / mm;
    for (uint32_t t = 0; t < T2; t++) batch_invert_body<F, M>(t, T2, NT, mm, tot.data());
    for (uint32_t t = 0; t < NT; t++) {
      if ((size_t)t * PB >= n_out) continue;
      if (chunked) pair_phase2<F, PB, true>(tot[t], srcg.data() + t, NT, cur_vals, cur, preg.data() + t, NT, outx.data(), outy.data(), dstg.data() + t);
      else if (l == 0) pair_phase2<F, PB, true>(tot[t], srcg.data() + t, NT, cur_vals, cur, preg.data() + t, NT, outx.data() + (size_t)t * PB, outy.data() + (size_t)t * PB);
      else pair_phase2<F, PB, false>(tot[t], srcg.data() + t, NT, nullptr, cur, preg.data() + t, NT, outx.data() + (size_t)t * PB, outy.data() + (size_t)t * PB);
    }
    if (chunked) {      // level-1 offsets of the bucket slots: where each slot's chunk-0 run was written
      next_off.assign(sh.nbuckets + 1, 0);
      for (uint32_t k = 0; k <= sh.nbuckets; k++) next_off[k] = dst_base[(size_t)k * sh.nchunks];
    }
    lvx.swap(outx);
    lvy.swap(outy);
    cur = level_from_xy<F>(lvx.data(), lvy.data());
    cur_vals = nullptr;
    offsets.swap(next_off);
  }
  std::vector<uint32_t> counts(sh.nbuckets), tstart(sh.nbuckets);
  for (uint32_t b = 0; b < sh.nbuckets; b++) msm_task_count_body(b, offsets.data(), sh.nbuckets, sh.Bs, K, counts.data());
  uint32_t ntasks = 0;
  for (uint32_t b = 0; b < sh.nbuckets; b++) { tstart[b] = ntasks; ntasks += counts[b]; }
  std::vector<uint2> tasks(ntasks ? ntasks : 1);
  for (uint32_t b = 0; b < sh.nbuckets; b++) msm_build_tasks_body(b, offsets.data(), tstart.data(), sh.nbuckets, sh.Bs, K, tasks.data());
  std::vector<xyzz_t<F>> partial(ntasks ? ntasks : 1), buckets(sh.nbuckets);
  for (uint32_t t = 0; t < ntasks; t++) msm_accumulate_body<F>(t, ntasks, tasks.data(), cur_vals, cur, partial.data());
  for (uint32_t b = 0; b < sh.nbuckets; b++) msm_finalize_body<F>(b, sh.nbuckets, tstart.data(), counts.data(), partial.data(), buckets.data());
  if (L > sh.B) L = sh.B;
  uint32_t segs = sh.B / L;
  const uint32_t WB = sh.Wf * batch;               // bucket sets of the whole batch
  std::vector<xyzz_t<F>> seg(WB * segs);
  for (uint32_t g = 0; g < WB * segs; g++) msm_segment_body<F>(g, WB, sh.B, L, buckets.data(), seg.data());
  for (uint32_t half = segs / 2; half >= 1; half /= 2)
    for (uint32_t g = 0; g < WB * half; g++) msm_tree_body<F>(g, WB, segs, half, seg.data());
  for (uint32_t b = 0; b < batch; b++) {
    xyzz_t<F> r = msm_combine<F>(seg.data() + (size_t)b * sh.Wf * segs, segs, sh.Wf, sh.c);
    jacobian_t<F> o = msm_result_encode<F>(r, false);
    const unsigned char* p = (const unsigned char*)&o;
    for (size_t i = 0; i < sizeof(o); i++) printf("%02x", p[i]);
    printf("\n");
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 8) return 1;
  bool g2 = argv[1][1] == '2';
  uint32_t n = atoi(argv[2]), c = atoi(argv[3]), K = atoi(argv[4]), L = atoi(argv[5]);
  bool mont = atoi(argv[6]) != 0;
  FILE* f = fopen(argv[7], "rb");
  if (!f) return 3;
  uint32_t factor = argc > 8 ? atoi(argv[8]) : 1;
  uint32_t levels = argc > 9 ? atoi(argv[9]) : 0;
  uint32_t batch = argc > 10 ? atoi(argv[10]) : 1;
  bool shared = argc > 11 ? atoi(argv[11]) != 0 : true;
  uint32_t chunk_log = argc > 12 ? atoi(argv[12]) : 31;
  uint32_t piece_chunks = argc > 13 ? atoi(argv[13]) : 0;    // > 0: level 0 streamed in pieces of that many chunks
  int rc = g2 ? run<fq2_t>(n, c, K, L, mont, f, factor, levels, batch, shared, chunk_log, piece_chunks)
              : run<fq_t>(n, c, K, L, mont, f, factor, levels, batch, shared, chunk_log, piece_chunks);
  fclose(f);
  return rc;
}
