// CPU-only driver for the per-thread NTT bodies in csrc/ntt_core.cuh (TEST INFRASTRUCTURE).
// Emulates k_ntt_pass (load / S butterfly stages / store, with barriers between phases) and the
// host-side pass planner + buffer routing of csrc/ntt.cu.  Usage:
//   ntt_host_sim <log_n> <batch> <inverse 0|1> <ordering 0..3> <columns 0|1> <coset 0|1> <inplace 0|1> <infile>
// infile: root (32 B, Montgomery, order 2^log_n) | coset_gen (32 B, Montgomery) | batch*N elements.
// prints batch*N output elements as hex.
#include <cstdio>
#include <cstdlib>
#include <vector>
#define B381_HOST_TEST 1
#include "ntt_core.cuh"
using namespace b381;

static void run_pass(const ntt_pass_params& p, const fr_t* in, fr_t* out) {
  const uint32_t tile_log = p.S + p.g + p.x, te = 1u << tile_log;
  const uint64_t tiles = (p.total + te - 1) >> tile_log;
  std::vector<uint4> lo(te), hi(te);
  ntt_tile t{lo.data(), hi.data()};
  for (uint64_t tile = 0; tile < tiles; tile++) {
    for (uint32_t pos = 0; pos < te; pos++) ntt_tile_load(p, tile, pos, in, t);
    if (tile & 1) {                        // odd tiles: stage by stage (the plain radix-2 body)
      for (int s = (int)p.S - 1; s >= 0; s--)
        for (uint32_t q = 0; q < te / 2; q++) ntt_tile_stage(p, tile, q, (uint32_t)s, t);
    } else {                               // even tiles: register-blocked steps exactly as k_ntt_pass runs them
      uint32_t s = p.S;
      while (s > 0) {
        const uint32_t R = s >= 3 ? 3u : s;
        s -= R;
        for (uint32_t q = 0; q < (te >> R); q++) {
          if (R == 3) ntt_tile_stages<3>(p, tile, q, s, t);
          else if (R == 2) ntt_tile_stages<2>(p, tile, q, s, t);
          else ntt_tile_stages<1>(p, tile, q, s, t);
        }
      }
    }
    for (uint32_t pos = 0; pos < te; pos++) ntt_tile_store(p, tile, pos, out, t);
  }
}

int main(int argc, char** argv) {
  if (argc < 9) return 1;
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
  std::vector<fr_t> table((1ull << K) - 1 + 1);
  {
    uint64_t half = 1ull << (K - 1);
    fr_t* top = table.data() + (half - 1);
    for (uint64_t c = 0; c * 64 < half; c++) fr_powers_chunk(c, 64, half, root, one<fr_t>(), top);
    for (uint64_t idx = 0; idx + 1 < half; idx++) {
      uint32_t k = 63 - __builtin_clzll(idx + 1);
      uint64_t j = idx + 1 - (1ull << k);
      table[idx] = top[j << (K - 1 - k)];
    }
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
  int P = ntt_plan_passes(n, plan);
  fr_t* work = d_out;
  if (P >= 2 && (perm_in || perm_out) && (d_in == d_out || perm_out)) work = scratch.data();
  for (int i = 0; i < P; i++) {
    ntt_pass_params p = {};
    p.n = n; p.lo = plan[i].lo; p.S = plan[i].S; p.g = plan[i].g; p.x = plan[i].x;
    p.total = total;
    if (columns) { p.estride = batch; p.bstride = 1; } else { p.estride = 1; p.bstride = N; }
    p.inverse = inverse; p.twiddles = table.data();
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
