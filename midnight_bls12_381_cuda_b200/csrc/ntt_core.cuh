// Per-thread bodies of the Fr NTT (tile load / butterfly stage / tile store), shared by the CUDA
// kernel in ntt.cu and the CPU-only harness tests/host/ntt_host_sim.cpp.
//
// Algorithm (ours; the reference's is bls12-381/src/field/ntt_kernels.cu:110-958):
//   * decimation-in-frequency radix-2 over log2(N) stages, grouped into PASSES of up to 11 stages;
//     a pass keeps a 2048-element tile (64 KB) in shared memory, so N = 2^24 costs 3 trips through
//     HBM instead of the reference's 12 radix-4 passes + bit-reversal pass + D2D copy (:772-810).
//   * stage k butterfly on (i, i+2^k):  (a, b) -> (a+b, (a-b) * w),  w = omega_{2^(k+1)}^(i mod 2^k).
//     Twiddles come from ONE stage-major table  T[2^k - 1 + j] = omega_{2^(k+1)}^j  (j < 2^k) that
//     serves every transform size of the domain (the reference keeps a full-length table per log
//     size, ~2 GiB for max-log 24: ntt_kernels.cu:1607-1679).  Inverse twiddles are read from the
//     same table:  w^-j = -T[2^k - j].
//   * DIF leaves y[bitrev(i)] at position i, so ordering is a pure addressing choice:
//     load position i from in[perm_in(i)], store position i to out[perm_out(i)]
//       kNN: in i      / out rev(i)        kNR: in i      / out i
//       kRN: in rev(i) / out rev(i)        kRR: in rev(i) / out i
//     (ICICLE Ordering enum, bls12-381/include/icicle_types.cuh:89-96; the reference's registered
//     path ignores `ordering`).
//   * coset (x[i]*g^i before a forward NTT, y[k]*g^-k after an inverse one; include/ntt.cuh:123-183)
//     and the 1/N of the inverse are fused into the first pass's load / last pass's store.
#pragma once
#include "field.cuh"

namespace b381 {

struct ntt_pass_params {
  uint32_t n;          // log2 of the transform length
  uint32_t lo;         // this pass runs stages k = lo+S-1 ... lo
  uint32_t S;          // stages in this pass
  uint32_t g;          // log2 of adjacent low-index elements kept together in a tile (coalescing)
  uint32_t x;          // lo == 0 only: log2 of whole sub-transforms packed into one tile (small N)
  uint64_t total;      // batch * N elements
  uint64_t estride;    // address = batch_index * bstride + i * estride
  uint64_t bstride;
  uint32_t inverse;    // use conjugate twiddles
  uint32_t perm_in;    // gather input from bit-reversed index (first pass only)
  uint32_t perm_out;   // scatter output to bit-reversed index (last pass only)
  const fr_t* twiddles;   // stage-major table
  const fr_t* pre_scale;  // optional: multiply position i by pre_scale[i] on load  (coset, forward)
  const fr_t* post_scale; // optional: multiply logical output k by post_scale[k] on store (coset, inverse)
  fr_t post_const;        // used when has_post_const: multiply every output (1/N)
  uint32_t has_post_const;
  // Distributed (multi-GPU, four-step) use: this GPU holds the column block l in [l_base, l_base+2^logL)
  // of a 2^n_glob transform as a local array J = i_hi * 2^logL + (l - l_base); the pass runs on the
  // local index space (n = n_glob - log2(#GPUs)) but twiddles follow the GLOBAL index.
  uint32_t dist_shift;    // log2(#GPUs); 0 = single-GPU transform
  uint32_t dist_logL;     // log2 of local columns
  uint32_t dist_lo;       // global bit position where the column index ends (= dist_logL + dist_shift)
  uint32_t dist_lbase;    // first global column owned by this GPU
  // Fused exchange (last column pass only): instead of the local array, element (row i, column l) is stored
  // straight into the ROW buffer of the GPU that owns row i -- peer memory over NVLink -- at
  // (i mod R) * 2^dist_lo + l_base + l, i.e. already in the [R][2^lo] layout the row transforms read.
  // The all_to_all and the transpose pass of the four-step NTT disappear.
  uint32_t peer_on;
  uint32_t peer_logR;     // log2 of rows per GPU
  uint32_t tile_rot;      // CTAs visit the tiles rotated by this much, so that at any moment the GPUs store to
                          // DIFFERENT peers (rank r starts with its own rows) instead of all hitting GPU 0 first
  fr_t* peer_out[8];      // row buffer of every GPU, mapped into this process (CUDA IPC)
};

// local flat index -> global flat index (identity when not distributed)
B381_DI uint64_t ntt_global_index(const ntt_pass_params& p, uint64_t J) {
  if (p.dist_shift == 0) return J;
  uint64_t hi = J >> p.dist_logL, l = J & ((1ull << p.dist_logL) - 1);
  return (hi << p.dist_lo) | (p.dist_lbase + l);
}

B381_HD uint32_t bitrev32(uint32_t v, uint32_t bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(v) >> (32 - bits)) : 0;
#else
  uint32_t r = 0;
  for (uint32_t i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

// shared-memory tile: element e lives as two 16-byte halves in separate arrays so that a warp
// touching consecutive elements hits all 32 banks exactly once per 8 lanes.  Slot `pos` is stored at
// pos ^ ((pos >> 3) & 7): a quarter-warp whose lanes step by 8 slots (the radix-8 step over the three
// lowest slot bits, one register group per lane) then still spreads over all eight 16-byte bank groups,
// and lanes on consecutive slots keep doing so (the XOR term is the same for all eight).
struct ntt_tile {
  uint4* lo;
  uint4* hi;
};
B381_HD uint32_t ntt_tile_slot(uint32_t pos) { return pos ^ ((pos >> 3) & 7u); }

B381_DI fr_t tile_get(const ntt_tile& t, uint32_t pos) {
  pos = ntt_tile_slot(pos);
  uint4 a = t.lo[pos], b = t.hi[pos];
  fr_t r;
  r.l[0] = ((uint64_t)a.y << 32) | a.x; r.l[1] = ((uint64_t)a.w << 32) | a.z;
  r.l[2] = ((uint64_t)b.y << 32) | b.x; r.l[3] = ((uint64_t)b.w << 32) | b.z;
  return r;
}
B381_DI void tile_put(const ntt_tile& t, uint32_t pos, const fr_t& v) {
  uint4 a, b;
  a.x = (uint32_t)v.l[0]; a.y = (uint32_t)(v.l[0] >> 32); a.z = (uint32_t)v.l[1]; a.w = (uint32_t)(v.l[1] >> 32);
  b.x = (uint32_t)v.l[2]; b.y = (uint32_t)(v.l[2] >> 32); b.z = (uint32_t)v.l[3]; b.w = (uint32_t)(v.l[3] >> 32);
  pos = ntt_tile_slot(pos);
  t.lo[pos] = a;
  t.hi[pos] = b;
}

B381_DI fr_t fr_gload(const fr_t* p) {
#if defined(__CUDA_ARCH__)
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  fr_t r;
  r.l[0] = ((uint64_t)a.y << 32) | a.x; r.l[1] = ((uint64_t)a.w << 32) | a.z;
  r.l[2] = ((uint64_t)b.y << 32) | b.x; r.l[3] = ((uint64_t)b.w << 32) | b.z;
  return r;
#else
  return *p;
#endif
}
B381_DI fr_t fr_gload_ro(const fr_t* p) {   // read-only path (twiddles, scale tables)
#if defined(__CUDA_ARCH__)
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fr_t r;
  r.l[0] = ((uint64_t)a.y << 32) | a.x; r.l[1] = ((uint64_t)a.w << 32) | a.z;
  r.l[2] = ((uint64_t)b.y << 32) | b.x; r.l[3] = ((uint64_t)b.w << 32) | b.z;
  return r;
#else
  return *p;
#endif
}
// one 256-bit store (STG.E.256, sm_100): the peer-memory stores of the fused exchange must not reach NVLink as two
// half-sector writes per element (measured at 8 GPUs: 16-byte stores made the column pass 2x slower)
B381_DI void fr_gstore256(fr_t* p, const fr_t& v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v.l[0]), "l"(v.l[1]), "l"(v.l[2]), "l"(v.l[3]) : "memory");
#else
  *p = v;
#endif
}
B381_DI void fr_gstore(fr_t* p, const fr_t& v) {
#if defined(__CUDA_ARCH__)
  uint4* q = reinterpret_cast<uint4*>(p);
  uint4 a, b;
  a.x = (uint32_t)v.l[0]; a.y = (uint32_t)(v.l[0] >> 32); a.z = (uint32_t)v.l[1]; a.w = (uint32_t)(v.l[1] >> 32);
  b.x = (uint32_t)v.l[2]; b.y = (uint32_t)(v.l[2] >> 32); b.z = (uint32_t)v.l[3]; b.w = (uint32_t)(v.l[3] >> 32);
  q[0] = a;
  q[1] = b;
#else
  *p = v;
#endif
}

// Flat index of tile slot `pos` (pos = m*G + l') of tile `tile_id`:
//   I = (H << hi) | (m << lo) | (l0 + l'),  hi = lo+S,  tiles enumerate (H, l0/G) with l0 fastest.
B381_DI uint64_t ntt_tile_index(const ntt_pass_params& p, uint64_t tile_id, uint32_t pos) {
  if (p.lo == 0) return (tile_id << (p.S + p.x)) | pos;   // contiguous tile (g == 0)
  const uint32_t G = 1u << p.g;
  const uint32_t groups_lo = (1u << p.lo) >> p.g;   // number of l0 groups (lo >= g by construction)
  uint64_t H = tile_id / groups_lo;
  uint32_t l0 = (uint32_t)(tile_id % groups_lo) << p.g;
  uint32_t m = pos >> p.g, lp = pos & (G - 1);
  return (H << (p.lo + p.S)) | ((uint64_t)m << p.lo) | (l0 + lp);
}

B381_DI uint64_t ntt_addr(const ntt_pass_params& p, uint64_t I, bool permute) {
  uint64_t b = I >> p.n;
  uint32_t i = (uint32_t)(I & ((1ull << p.n) - 1));
  if (permute) i = bitrev32(i, p.n);
  return b * p.bstride + (uint64_t)i * p.estride;
}

// phase 1: slot `pos` <- global
B381_DI void ntt_tile_load(const ntt_pass_params& p, uint64_t tile_id, uint32_t pos, const fr_t* in, const ntt_tile& t) {
  uint64_t I = ntt_tile_index(p, tile_id, pos);
  if (I >= p.total) return;
  fr_t v = fr_gload(in + ntt_addr(p, I, p.perm_in != 0));
  if (p.pre_scale) v = mul(v, fr_gload_ro(p.pre_scale + (I & ((1ull << p.n) - 1))));
  tile_put(t, pos, v);
}

// phase 2: butterfly q of in-pass stage s (global stage k = lo + s)
B381_DI void ntt_tile_stage(const ntt_pass_params& p, uint64_t tile_id, uint32_t q, uint32_t s, const ntt_tile& t) {
  const uint32_t bit = s + p.g;                       // slot bit that separates the pair
  uint32_t pos0 = ((q >> bit) << (bit + 1)) | (q & ((1u << bit) - 1));
  uint32_t pos1 = pos0 | (1u << bit);
  uint64_t I0 = ntt_tile_index(p, tile_id, pos0);
  if (I0 >= p.total) return;
  const uint32_t k = p.lo + s + p.dist_shift;         // global stage
  uint32_t j = (uint32_t)(ntt_global_index(p, I0) & ((1ull << k) - 1));    // i mod 2^k
  fr_t a = tile_get(t, pos0), b = tile_get(t, pos1);
  fr_t sum = add(a, b);
  fr_t d, w;
  const fr_t* T = p.twiddles + ((1ull << k) - 1);
  if (!p.inverse || j == 0) {
    d = sub(a, b);
    w = fr_gload_ro(T + j);
  } else {                                            // w^-j = -T[2^k - j]
    d = sub(b, a);
    w = fr_gload_ro(T + ((1u << k) - j));
  }
  tile_put(t, pos0, sum);
  tile_put(t, pos1, (k == 0) ? d : mul(d, w));        // stage 0 twiddle is 1
}

// One butterfly on registers, global stage k = lo + s (+ dist_shift), first element at tile slot pos0.
B381_DI void ntt_bfly(const ntt_pass_params& p, uint64_t tile_id, uint32_t pos0, uint32_t s, fr_t& a, fr_t& b) {
  const uint32_t k = p.lo + s + p.dist_shift;
  const uint64_t I0 = ntt_tile_index(p, tile_id, pos0);
  const uint32_t j = (uint32_t)(ntt_global_index(p, I0) & ((1ull << k) - 1));
  const fr_t sum = add(a, b);
  fr_t d, w;
  const fr_t* T = p.twiddles + ((1ull << k) - 1);
  if (!p.inverse || j == 0) {
    d = sub(a, b);
    w = fr_gload_ro(T + j);
  } else {
    d = sub(b, a);
    w = fr_gload_ro(T + ((1u << k) - j));
  }
  a = sum;
  b = (k == 0) ? d : mul(d, w);
}

// phase 2, register-blocked: R consecutive in-pass stages s0+R-1 .. s0 on the 2^R tile slots that differ
// in slot bits [s0+g, s0+g+R) -- 2^R elements travel smem -> registers -> smem once for R stages
// (R * 2^(R-1) butterflies), instead of once per stage.  Group q of the tile's 2^(S+g+x-R) groups.
template <int R>
B381_DI void ntt_tile_stages(const ntt_pass_params& p, uint64_t tile_id, uint32_t q, uint32_t s0, const ntt_tile& t) {
  const uint32_t bit0 = s0 + p.g;
  const uint32_t base = ((q >> bit0) << (bit0 + R)) | (q & ((1u << bit0) - 1));
  if (ntt_tile_index(p, tile_id, base) >= p.total) return;   // a group never straddles two transforms
  fr_t v[1 << R];
#pragma unroll
  for (int a = 0; a < (1 << R); a++) v[a] = tile_get(t, base | ((uint32_t)a << bit0));
#pragma unroll
  for (int r = R - 1; r >= 0; r--) {
#pragma unroll
    for (int a = 0; a < (1 << R); a++) {
      if (a & (1 << r)) continue;
      ntt_bfly(p, tile_id, base | ((uint32_t)a << bit0), s0 + (uint32_t)r, v[a], v[a | (1 << r)]);
    }
  }
#pragma unroll
  for (int a = 0; a < (1 << R); a++) tile_put(t, base | ((uint32_t)a << bit0), v[a]);
}

// phase 3: slot `pos` -> global
B381_DI void ntt_tile_store(const ntt_pass_params& p, uint64_t tile_id, uint32_t pos, fr_t* out, const ntt_tile& t) {
  uint64_t I = ntt_tile_index(p, tile_id, pos);
  if (I >= p.total) return;
  fr_t v = tile_get(t, pos);
  if (p.post_scale) {
    // position i holds logical output rev(i)
    uint32_t i = (uint32_t)(I & ((1ull << p.n) - 1));
    v = mul(v, fr_gload_ro(p.post_scale + bitrev32(i, p.n)));
  } else if (p.has_post_const) {
    v = mul(v, p.post_const);
  }
  if (p.peer_on) {
    const uint64_t row = I >> p.dist_logL;
    const uint32_t l = (uint32_t)(I & ((1ull << p.dist_logL) - 1));
    fr_t* dst = p.peer_out[row >> p.peer_logR];
    fr_gstore256(dst + (((row & ((1ull << p.peer_logR) - 1)) << p.dist_lo) | (p.dist_lbase + l)), v);
    return;
  }
  fr_gstore(out + ntt_addr(p, I, p.perm_out != 0), v);
}

// ---- domain / table generation bodies -------------------------------------------------------
// out[j] = base * g^j for j in [chunk*len, chunk*len+len) : one pow + (len-1) multiplications.
B381_DI void fr_powers_chunk(uint64_t chunk, uint32_t len, uint64_t count, const fr_t& g, const fr_t& base, fr_t* out) {
  uint64_t j0 = chunk * len;
  if (j0 >= count) return;
  fr_t cur = mul(base, pow_u64(g, j0));
  for (uint32_t t = 0; t < len && j0 + t < count; t++) {
    out[j0 + t] = cur;
    cur = mul(cur, g);
  }
}

// ---- pass planner (host) --------------------------------------------------------------------
constexpr uint32_t kNttTileLog = 11;   // 2048 elements = 64 KB of shared memory per CTA
struct ntt_pass_plan { uint32_t lo, S, g, x; };

// Last pass: up to 11 stages on contiguous tiles.  Earlier passes: at most 9 stages each, so that a
// tile row keeps >= 4 adjacent elements (128 B contiguous) for coalescing.
// Fills `out` (capacity 8) and returns the number of passes.
inline int ntt_plan_passes(uint32_t n, ntt_pass_plan* out) {
  int np = 0;
  if (n <= kNttTileLog) {
    out[np++] = ntt_pass_plan{0, n, 0, kNttTileLog - n};
    return np;
  }
  uint32_t rest = n - kNttTileLog;
  uint32_t upper = (rest + 8) / 9;
  uint32_t hi = n;
  for (uint32_t i = 0; i < upper; i++) {
    uint32_t S = (rest + (upper - i) - 1) / (upper - i);
    out[np++] = ntt_pass_plan{hi - S, S, kNttTileLog - S, 0};
    hi -= S;
    rest -= S;
  }
  out[np++] = ntt_pass_plan{0, kNttTileLog, 0, 0};
  return np;
}

}  // namespace b381
