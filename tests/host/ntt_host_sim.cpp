// CPU-only driver for the per-thread NTT bodies in csrc/ntt_core.cuh (TEST INFRASTRUCTURE).
// Emulates k_ntt_pass (load / S butterfly stages / store, with barriers between phases) and the
// host-side pass planner + buffer routing of csrc/ntt.cu.  Usage:
//   ntt_host_sim <log_n> <batch> <inverse 0|1> <ordering 0..3> <columns 0|1> <coset 0|1> <inplace 0|1> <infile> [tile_log]
// tile_log (default 11, the kernels' value) shrinks the tile so that small transforms exercise 3- and 4-pass plans.
// infile: root (32 B, Montgomery, order 2^log_n) | coset_gen (32 B, Montgomery) | batch*N elements.
// prints batch*N output elements as hex.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#define B381_HOST_TEST 1
#include "ntt_core.cuh"
using namespace b381;

// one pass exactly as k_ntt / k_ntt_generic run it: every "thread" (slot / group index) of a phase, then the barrier
static void run_pass(const ntt_pass_params& p, const fr_t* in, fr_t* out) {
  const uint32_t tile_log = p.S + p.g + p.x, te = 1u << tile_log;
  const uint64_t tiles = (p.total + te - 1) >> tile_log;
  std::vector<uint4> lo(te), hi(te);
  ntt_tile t{lo.data(), hi.data()};
  const bool last_is_unit = p.lo == 0 && p.dist_shift == 0;
  for (uint64_t tile = 0; tile < tiles; tile++) {
    const ntt_tile_ctx c = ntt_tile_begin(p, tile, p.S, p.g, p.x, p.lo);
    for (uint32_t pos = 0; pos < te; pos++) ntt_tile_load(p, c, pos, in, t);
    for (uint32_t step = 0; step < ntt_pass_steps(p.S); step++) {
      uint32_t rl, s;
      ntt_pass_step(p.S, step, &rl, &s);
      if (rl == 1) for (uint32_t q = 0; q < te / 2; q++) ntt_step_r2(p, c, q, s, t);
      else if (last_is_unit && s == 0) for (uint32_t q = 0; q < te / 4; q++) ntt_step_r4<true>(p, c, q, s, t);
      else for (uint32_t q = 0; q < te / 4; q++) ntt_step_r4<false>(p, c, q, s, t);
    }
    for (uint32_t pos = 0; pos < te; pos++) ntt_tile_store(p, c, pos, out, t);
  }
}

// Distributed (four-step) transform with every rank emulated in turn (csrc/ntt.cu ntt_dist_columns + dist.py):
//   ntt_host_sim dist <log_n> <log_gpus> <upper_stages a> <inverse> <fused 0|1> <infile> [tile_log]
// infile as below (batch = 1).  Each rank runs the column passes on its column block; the rows are exchanged (fused:
// the last pass stores straight into the owners' row buffers; else a plain transpose of row blocks); every rank runs
// its row transforms in kNR order.  Prints the concatenated row blocks = the global kNR result.
static int run_dist(int argc, char** argv) {
  if (argc < 8) return 1;
  const uint32_t n = atoi(argv[2]), lg = atoi(argv[3]), a = atoi(argv[4]);
  const bool inverse = atoi(argv[5]), fused = atoi(argv[6]);
  const uint32_t tl = argc > 8 ? (uint32_t)atoi(argv[8]) : kNttTileLog;
  FILE* f = fopen(argv[7], "rb");
  if (!f) return 2;
  fr_t root, gen;
  const uint64_t N = 1ull << n;
  std::vector<fr_t> x(N);
  if (fread(&root, 32, 1, f) != 1 || fread(&gen, 32, 1, f) != 1 || fread(x.data(), 32, N, f) != N) return 3;
  fclose(f);
  const uint64_t entries = N - 1, half = N / 2;
  std::vector<fr_t> table(2 * entries + 1);
  fr_t* top = table.data() + (half - 1);
  for (uint64_t c = 0; c * 64 < half; c++) fr_powers_chunk(c, 64, half, root, one<fr_t>(), top);
  for (uint64_t idx = 0; idx + 1 < half; idx++) {
    uint32_t k = 63 - __builtin_clzll(idx + 1);
    table[idx] = top[(idx + 1 - (1ull << k)) << (n - 1 - k)];
  }
  for (uint64_t idx = 0; idx < entries; idx++) ntt_inverse_twiddle(idx, table.data(), table.data() + entries);
  const fr_t* tw = inverse ? table.data() + entries : table.data();
  fr_t ninv = one<fr_t>(), half_e = inv(add(one<fr_t>(), one<fr_t>()));
  const uint32_t G = 1u << lg, lo = n - a, logL = lo - lg;
  const uint64_t L = 1ull << logL, rows = 1ull << a, R = rows / G, loc = N / G;
  std::vector<std::vector<fr_t>> rowbuf(G, std::vector<fr_t>(loc)), colbuf(G, std::vector<fr_t>(loc));
  for (uint32_t r = 0; r < G; r++) {
    for (uint64_t i = 0; i < rows; i++)
      for (uint64_t l = 0; l < L; l++) colbuf[r][i * L + l] = x[(i << lo) + r * L + l];
    ntt_pass_params passes[8];
    const int np = ntt_dist_passes(n, lg, r, a, tw, passes, tl);
    for (int i = 0; i < np; i++) {
      ntt_pass_params& p = passes[i];
      if (fused && i + 1 == np) {
        p.peer_on = 1;
        p.peer_logR = a - lg;
        for (uint32_t d = 0; d < G; d++) p.peer_out[d] = rowbuf[d].data();
      }
      run_pass(p, colbuf[r].data(), colbuf[r].data());
    }
    if (!fused)   // all_to_all + transpose: row i of this rank's columns goes to rank i / R at (i mod R) * 2^lo + r*L + l
      for (uint64_t i = 0; i < rows; i++)
        for (uint64_t l = 0; l < L; l++) rowbuf[i / R][((i % R) << lo) + r * L + l] = colbuf[r][i * L + l];
  }
  for (uint32_t i = 0; i < lo; i++) ninv = mul(ninv, half_e);   // the row transforms scale by 2^-lo ...
  fr_t ninv_a = one<fr_t>();
  for (uint32_t i = 0; i < a; i++) ninv_a = mul(ninv_a, half_e);   // ... and the caller owes 2^-a (dist.py applies it)
  for (uint32_t r = 0; r < G; r++) {
    ntt_pass_plan plan[8];
    const int P = ntt_plan_passes(lo, plan, tl);
    for (int i = 0; i < P; i++) {
      ntt_pass_params p = {};
      p.n = lo; p.lo = plan[i].lo; p.S = plan[i].S; p.g = plan[i].g; p.x = plan[i].x;
      p.total = loc; p.estride = 1; p.bstride = 1ull << lo;
      p.twiddles = tw;
      if (i + 1 == P && inverse) { p.post_const = mul(ninv, ninv_a); p.has_post_const = 1; }
      run_pass(p, rowbuf[r].data(), rowbuf[r].data());
    }
    const unsigned char* o = (const unsigned char*)rowbuf[r].data();
    for (uint64_t i = 0; i < loc * 32; i++) printf("%02x", o[i]);
  }
  printf("\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "dist") return run_dist(argc, argv);
  if (argc < 9) return 1;
  const uint32_t tl = argc > 9 ? (uint32_t)atoi(argv[9]) : kNttTileLog;
  uint32_t n = atoi(argv[1]), batch = atoi(argv[2]);
  bool inverse = atoi(argv[3]), columns = atoi(argv[5]), coset = atoi(argv[6]), inplace = atoi(argv[7]);
  int ordering = atoi(argv[4]);
  FILE* f = fopen(argv[8], "rb");
  if (!f) return 2;
  fr_t root, gen;
  uint64_t N = 1ull << n, total = N * batch;
  std::vector<fr_t> data(total);
  if (fread(&root, 32, 1, f) != 1 || fread(&gen, 32, 1, f) != 1 || fread(data.data(), 32, total, f) != total) return 3;
  fclose(f);
  // domain: stage-major table for K = n (same construction as init_domain in ntt.cu)
  uint32_t K = n ? n : 1;
  const uint64_t entries = (1ull << K) - 1;
  std::vector<fr_t> table(2 * entries + 1);
  {
    uint64_t half = 1ull << (K - 1);
    fr_t* top = table.data() + (half - 1);
    for (uint64_t c = 0; c * 64 < half; c++) fr_powers_chunk(c, 64, half, root, one<fr_t>(), top);
    for (uint64_t idx = 0; idx + 1 < half; idx++) {
      uint32_t k = 63 - __builtin_clzll(idx + 1);
      uint64_t j = idx + 1 - (1ull << k);
      table[idx] = top[j << (K - 1 - k)];
    }
    for (uint64_t idx = 0; idx < entries; idx++) ntt_inverse_twiddle(idx, table.data(), table.data() + entries);
  }
  fr_t ninv = one<fr_t>(), half_e = inv(add(one<fr_t>(), one<fr_t>()));
  for (uint32_t i = 0; i < n; i++) ninv = mul(ninv, half_e);
  std::vector<fr_t> ctab;
  if (coset) {
    ctab.resize(N);
    fr_t g = inverse ? inv(gen) : gen, base = inverse ? ninv : one<fr_t>();
    for (uint64_t c = 0; c * 64 < N; c++) fr_powers_chunk(c, 64, N, g, base, ctab.data());
  }
  bool perm_in = ordering == 2 || ordering == 3, perm_out = ordering == 0 || ordering == 2;
  std::vector<fr_t> outbuf(total), scratch(total);
  const fr_t* d_in = data.data();
  fr_t* d_out = inplace ? data.data() : outbuf.data();
  ntt_pass_plan plan[8];
  int P = ntt_plan_passes(n, plan, tl);
  fr_t* work = d_out;
  if (P >= 2 && (perm_in || perm_out) && (d_in == d_out || perm_out)) work = scratch.data();
  for (int i = 0; i < P; i++) {
    ntt_pass_params p = {};
    p.n = n; p.lo = plan[i].lo; p.S = plan[i].S; p.g = plan[i].g; p.x = plan[i].x;
    p.total = total;
    if (columns) { p.estride = batch; p.bstride = 1; } else { p.estride = 1; p.bstride = N; }
    p.twiddles = inverse ? table.data() + entries : table.data();
    bool first = i == 0, last = i + 1 == P;
    p.perm_in = first && perm_in; p.perm_out = last && perm_out;
    if (first && coset && !inverse) p.pre_scale = ctab.data();
    if (last && inverse) { if (coset) p.post_scale = ctab.data(); else { p.post_const = ninv; p.has_post_const = 1; } }
    run_pass(p, first ? d_in : work, last ? d_out : work);
  }
  const unsigned char* o = (const unsigned char*)d_out;
  for (uint64_t i = 0; i < total * 32; i++) printf("%02x", o[i]);
  printf("\n");
  return 0;
}
