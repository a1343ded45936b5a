// Per-thread bodies of the Fr NTT (tile load / register-blocked butterfly steps / tile store) and the host-side pass
// planner, shared by the CUDA kernels in ntt.cu and the CPU-only harness tests/host/ntt_host_sim.cpp.
//
// Algorithm (ours; the reference's is bls12-381/src/field/ntt_kernels.cu:110-958):
//   * decimation-in-frequency radix-2 over log2(N) stages, grouped into PASSES of S stages; a pass keeps a
//     2^TL-element tile in shared memory (TL = 10: 32 KB), so N = 2^24 is 3 trips through HBM (8 + 8 + 8 stages)
//     instead of the reference's 12 radix-4 passes + bit-reversal pass + D2D copy (:772-810).
//   * stage k butterfly on (i, i+2^k):  (a, b) -> (a+b, (a-b) * w),  w = omega_{2^(k+1)}^(i mod 2^k).
//     Inside a pass two stages are run at a time on FOUR tile slots held in registers (a radix-4 step: 4 butterflies,
//     3 distinct twiddles, one trip smem -> registers -> smem per two stages); the step is straight-line code -- no
//     data- or parameter-dependent branch between the loads and the stores -- so the two independent Montgomery
//     products of every stage interleave.  An odd S starts with one plain radix-2 step.
//   * Twiddles come from stage-major tables  T[2^k - 1 + j] = omega_{2^(k+1)}^j  (j < 2^k), one for omega and one for
//     omega^-1 (the inverse transform is the same code with the other table), which serve every transform size of the
//     domain (the reference keeps a full-length table per log size, ~2 GiB for max-log 24: ntt_kernels.cu:1607-1679).
//     In the last pass the final step (stages 1 and 0) needs one product per four elements: T_0[0] = T_1[0] = 1.
//   * DIF leaves y[bitrev(i)] at position i, so ordering is a pure addressing choice:
//     load position i from in[perm_in(i)], store position i to out[perm_out(i)]
//       kNN: in i      / out rev(i)        kNR: in i      / out i
//       kRN: in rev(i) / out rev(i)        kRR: in rev(i) / out i
//     (ICICLE Ordering enum, bls12-381/include/icicle_types.cuh:89-96; the reference's registered
//     path ignores `ordering`).
//   * coset (x[i]*g^i before a forward NTT, y[k]*g^-k after an inverse one; include/ntt.cuh:123-183)
//     and the 1/N of the inverse are fused into the first pass's load / last pass's store.
#pragma once
#include <cstring>

#include "field.cuh"

namespace b381 {

struct ntt_pass_params {
  uint32_t n;          // log2 of the transform length
  uint32_t lo;         // this pass runs stages k = lo+S-1 ... lo
  uint32_t S;          // stages in this pass
  uint32_t g;          // log2 of adjacent low-index elements kept together in a tile (coalescing); tile = 2^(S+g+x)
  uint32_t x;          // lo == 0 only: log2 of whole 2^S-blocks packed into one tile
  uint64_t total;      // batch * N elements
  uint64_t estride;    // address = batch_index * bstride + i * estride
  uint64_t bstride;
  uint32_t perm_in;    // gather input from bit-reversed index (first pass only)
  uint32_t perm_out;   // scatter output to bit-reversed index (last pass only)
  const fr_t* twiddles;   // stage-major table of omega (forward) or omega^-1 (inverse)
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
// pos ^ ((pos >> 3) & 7): a quarter-warp whose lanes step by 8 slots then still spreads over all eight 16-byte bank
// groups, and lanes on consecutive slots keep doing so (the XOR term is the same for all eight).
struct ntt_tile {
  uint4* lo;
  uint4* hi;
};
B381_HD uint32_t ntt_tile_slot(uint32_t pos) { return pos ^ ((pos >> 3) & 7u); }

// the 16-byte halves travel as two 64-bit words (ulonglong2 views of the same storage): an fr_t limb IS such a word,
// so no 32-bit halves have to be packed or unpacked around a load or a store
B381_DI fr_t tile_get(const ntt_tile& t, uint32_t pos) {
  pos = ntt_tile_slot(pos);
#if defined(__CUDA_ARCH__)
  const ulonglong2 a = reinterpret_cast<const ulonglong2*>(t.lo)[pos], b = reinterpret_cast<const ulonglong2*>(t.hi)[pos];
  fr_t r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = b.x; r.l[3] = b.y;
  return r;
#else
  uint4 a = t.lo[pos], b = t.hi[pos];
  fr_t r;
  r.l[0] = ((uint64_t)a.y << 32) | a.x; r.l[1] = ((uint64_t)a.w << 32) | a.z;
  r.l[2] = ((uint64_t)b.y << 32) | b.x; r.l[3] = ((uint64_t)b.w << 32) | b.z;
  return r;
#endif
}
B381_DI void tile_put(const ntt_tile& t, uint32_t pos, const fr_t& v) {
#if defined(__CUDA_ARCH__)
  pos = ntt_tile_slot(pos);
  reinterpret_cast<ulonglong2*>(t.lo)[pos] = make_ulonglong2(v.l[0], v.l[1]);
  reinterpret_cast<ulonglong2*>(t.hi)[pos] = make_ulonglong2(v.l[2], v.l[3]);
#else
  uint4 a, b;
  a.x = (uint32_t)v.l[0]; a.y = (uint32_t)(v.l[0] >> 32); a.z = (uint32_t)v.l[1]; a.w = (uint32_t)(v.l[1] >> 32);
  b.x = (uint32_t)v.l[2]; b.y = (uint32_t)(v.l[2] >> 32); b.z = (uint32_t)v.l[3]; b.w = (uint32_t)(v.l[3] >> 32);
  pos = ntt_tile_slot(pos);
  t.lo[pos] = a;
  t.hi[pos] = b;
#endif
}

B381_DI fr_t fr_gload(const fr_t* p) {
#if defined(__CUDA_ARCH__)
  const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
  const ulonglong2 a = q[0], b = q[1];
  fr_t r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = b.x; r.l[3] = b.y;
  return r;
#else
  return *p;
#endif
}
B381_DI fr_t fr_gload_ro(const fr_t* p) {   // read-only path (twiddles, scale tables)
#if defined(__CUDA_ARCH__)
  const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
  const ulonglong2 a = __ldg(q), b = __ldg(q + 1);
  fr_t r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = b.x; r.l[3] = b.y;
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
  ulonglong2* q = reinterpret_cast<ulonglong2*>(p);
  q[0] = make_ulonglong2(v.l[0], v.l[1]);
  q[1] = make_ulonglong2(v.l[2], v.l[3]);
#else
  *p = v;
#endif
}

// ---- what a CTA knows about its tile, computed once (uniform across the CTA) --------------------------------
// Flat index of tile slot `pos` = (m << g) | l'  (m: the S + x bits the pass works on, l': g adjacent low bits):
//   I = base + (m << lo) + l',   base = (H << (lo + S)) | l0   with tiles enumerating (H, l0 >> g), l0 fastest.
// (lo == 0 has g == 0 and I = (tile << (S + x)) | pos, the same formula with base = tile << (S + x).)
// The pass shape (S, g, x, lo) travels in here rather than being read from the parameter block, so that a kernel
// specialised on the shape hands the steps compile-time constants.
struct ntt_tile_ctx {
  uint32_t S, g, x, lo; // shape of the pass (see ntt_pass_params)
  uint64_t base;        // flat index of slot 0
  uint32_t lglob;       // GLOBAL low index part of slot 0's column l0 (== l0 when not distributed)
  uint32_t jshift;      // bit position of m in the GLOBAL index: lo (+ dist_shift)
};
B381_DI ntt_tile_ctx ntt_tile_begin(const ntt_pass_params& p, uint64_t tile_id, uint32_t S, uint32_t g, uint32_t x, uint32_t lo) {
  ntt_tile_ctx c;
  c.S = S; c.g = g; c.x = x; c.lo = lo;
  uint32_t l0 = 0;
  if (lo == 0) {
    c.base = tile_id << (S + x);
  } else {
    const uint32_t glog = lo - g;                         // log2 of l0 groups per H (lo >= g by construction)
    const uint64_t H = tile_id >> glog;
    l0 = (uint32_t)(tile_id & ((1ull << glog) - 1)) << g;
    c.base = (H << (lo + S)) | l0;
  }
  c.jshift = lo + (lo ? p.dist_shift : 0u);           // distributed passes never have lo == 0
  c.lglob = l0;
  if (p.dist_shift) {
    // local low part l0 spans column bits [0, logL) and row bits [logL, lo): the rows move up by dist_shift and the
    // column gets this GPU's first column added.  Slots add l' < 2^g <= 2^logL to the column part only.
    const uint32_t col = l0 & ((1u << p.dist_logL) - 1), row = l0 >> p.dist_logL;
    c.lglob = (row << p.dist_lo) | (p.dist_lbase + col);
  }
  return c;
}
B381_DI uint64_t ntt_slot_index(const ntt_tile_ctx& c, uint32_t pos) {
  return c.base + ((uint64_t)(pos >> c.g) << c.lo) + (pos & ((1u << c.g) - 1));
}

B381_DI uint64_t ntt_addr(const ntt_pass_params& p, uint64_t I, bool permute) {
  uint64_t b = I >> p.n;
  uint32_t i = (uint32_t)(I & ((1ull << p.n) - 1));
  if (permute) i = bitrev32(i, p.n);
  return b * p.bstride + (uint64_t)i * p.estride;
}

// 16-byte asynchronous copy global -> shared (LDGSTS): the tile load of a pass without a load-time scale keeps no
// registers and no scoreboard slot per element, so every element of the thread is in flight at once
B381_DI void cp_async16(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
#else
  memcpy(smem_dst, gsrc, 16);
#endif
}
B381_DI void cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
}

// phase 1: slot `pos` <- global  (followed by cp_async_wait_all() and the CTA barrier)
B381_DI void ntt_tile_load(const ntt_pass_params& p, const ntt_tile_ctx& c, uint32_t pos, const fr_t* in, const ntt_tile& t) {
  uint64_t I = ntt_slot_index(c, pos);
  if (I >= p.total) return;
  if (!p.pre_scale) {
    const char* src = reinterpret_cast<const char*>(in + ntt_addr(p, I, p.perm_in != 0));
    const uint32_t slot = ntt_tile_slot(pos);
    cp_async16(t.lo + slot, src);
    cp_async16(t.hi + slot, src + 16);
    return;
  }
  fr_t v = fr_gload(in + ntt_addr(p, I, p.perm_in != 0));
  if (p.pre_scale) v = mul(v, fr_gload_ro(p.pre_scale + (I & ((1ull << p.n) - 1))));
  tile_put(t, pos, v);
}

// (a, b) -> (a + b, (a - b) * w)
B381_DI void ntt_bfly(fr_t& a, fr_t& b, const fr_t& w) {
  const fr_t d = sub(a, b);
  a = add(a, b);
  b = mul(d, w);
}
B381_DI void ntt_bfly_unit(fr_t& a, fr_t& b) {   // twiddle 1
  const fr_t d = sub(a, b);
  a = add(a, b);
  b = d;
}

// twiddle index of the group whose first slot is `base_pos` (slot bits [bit0, bit0 + R) clear), in-pass stage s:
//   j0 = ((m mod 2^s) << jshift) + lglob + l'       (the GLOBAL index of the slot, modulo 2^(jshift + s))
B381_DI uint32_t ntt_twiddle_j0(const ntt_tile_ctx& c, uint32_t base_pos, uint32_t s) {
  const uint32_t m = base_pos >> c.g, lp = base_pos & ((1u << c.g) - 1);
  return ((m & ((1u << s) - 1)) << c.jshift) + c.lglob + lp;
}

// phase 2a: one radix-2 step (in-pass stage s) on group q of the tile's 2^(TL-1) pairs
B381_DI void ntt_step_r2(const ntt_pass_params& p, const ntt_tile_ctx& c, uint32_t q, uint32_t s, const ntt_tile& t) {
  const uint32_t bit0 = s + c.g;
  const uint32_t pos0 = ((q >> bit0) << (bit0 + 1)) | (q & ((1u << bit0) - 1)), pos1 = pos0 | (1u << bit0);
  const uint32_t k = c.jshift + s;
  const fr_t w = fr_gload_ro(p.twiddles + ((1ull << k) - 1) + ntt_twiddle_j0(c, pos0, s));
  fr_t a = tile_get(t, pos0), b = tile_get(t, pos1);
  ntt_bfly(a, b, w);
  tile_put(t, pos0, a);
  tile_put(t, pos1, b);
}

B381_DI void prefetch_l1(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// pull the three twiddles of radix-4 group q (in-pass stages s+1, s) towards the SM ahead of their use
B381_DI void ntt_prefetch_r4(const ntt_pass_params& p, const ntt_tile_ctx& c, uint32_t q, uint32_t s) {
  const uint32_t bit0 = s + c.g;
  const uint32_t base = ((q >> bit0) << (bit0 + 2)) | (q & ((1u << bit0) - 1));
  const uint32_t k = c.jshift + s;
  const uint32_t j0 = ntt_twiddle_j0(c, base, s);
  const fr_t* Tk1 = p.twiddles + ((2ull << k) - 1) + j0;
  prefetch_l1(p.twiddles + ((1ull << k) - 1) + j0);
  prefetch_l1(Tk1);
  prefetch_l1(Tk1 + (1ull << k));
}

// phase 2b: one radix-4 step = in-pass stages s+1 and s on group q of the tile's 2^(TL-2) quadruples: slots
// base | a << bit0, a = 0..3.  Stage s+1 pairs (0,2) and (1,3) with twiddles T_{k+1}[j0] and T_{k+1}[j0 + 2^k]; stage s
// pairs (0,1) and (2,3), both with T_k[j0]  (k = jshift + s: global stage of in-pass stage s).
// LAST = the pass's lo is 0 and s == 0: T_0[0] = T_1[0] = 1, only T_1[1] = omega_4 is a real product.
template <bool LAST>
B381_DI void ntt_step_r4(const ntt_pass_params& p, const ntt_tile_ctx& c, uint32_t q, uint32_t s, const ntt_tile& t) {
  const uint32_t bit0 = s + c.g;
  const uint32_t base = ((q >> bit0) << (bit0 + 2)) | (q & ((1u << bit0) - 1));
  const uint32_t p1 = base | (1u << bit0), p2 = base | (2u << bit0), p3 = base | (3u << bit0);
  if (LAST) {
    const fr_t w4 = fr_gload_ro(p.twiddles + 2);          // T_1[1] = omega_4 (or its inverse)
    fr_t x0 = tile_get(t, base), x1 = tile_get(t, p1), x2 = tile_get(t, p2), x3 = tile_get(t, p3);
    ntt_bfly_unit(x0, x2);
    ntt_bfly(x1, x3, w4);
    ntt_bfly_unit(x0, x1);
    ntt_bfly_unit(x2, x3);
    tile_put(t, base, x0); tile_put(t, p1, x1); tile_put(t, p2, x2); tile_put(t, p3, x3);
    return;
  }
  const uint32_t k = c.jshift + s;
  const uint32_t j0 = ntt_twiddle_j0(c, base, s);
  const fr_t* Tk = p.twiddles + ((1ull << k) - 1) + j0;
  const fr_t* Tk1 = p.twiddles + ((2ull << k) - 1) + j0;
  const fr_t wa = fr_gload_ro(Tk1), wb = fr_gload_ro(Tk1 + (1ull << k)), wl = fr_gload_ro(Tk);
  fr_t x0 = tile_get(t, base), x1 = tile_get(t, p1), x2 = tile_get(t, p2), x3 = tile_get(t, p3);
  ntt_bfly(x0, x2, wa);
  ntt_bfly(x1, x3, wb);
  ntt_bfly(x0, x1, wl);
  ntt_bfly(x2, x3, wl);
  tile_put(t, base, x0); tile_put(t, p1, x1); tile_put(t, p2, x2); tile_put(t, p3, x3);
}

// phase 3: slot `pos` -> global
B381_DI void ntt_tile_store(const ntt_pass_params& p, const ntt_tile_ctx& c, uint32_t pos, fr_t* out, const ntt_tile& t) {
  uint64_t I = ntt_slot_index(c, pos);
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

// The whole pass as seen by thread `tid` of `nthreads`, with `sync()` the CTA barrier: exactly what the CUDA kernel
// runs (ntt.cu) and what the host harness replays thread by thread, phase by phase.
// TL = log2 of the tile; S = p.S is a template parameter so that every step's shifts and masks are constants.
template <int TL>
struct ntt_pass_shape {
  static constexpr uint32_t tile = 1u << TL;
};
// phase list of a pass with S stages: optional radix-2 step at stage S-1 (S odd), then radix-4 steps down to stage 0.
// step index -> (radix, s): used identically by the kernel's unrolled loop and by the host harness.
B381_HD uint32_t ntt_pass_steps(uint32_t S) { return (S + 1) / 2; }
B381_HD void ntt_pass_step(uint32_t S, uint32_t step, uint32_t* radix_log, uint32_t* s) {
  if (S & 1) {
    if (step == 0) { *radix_log = 1; *s = S - 1; return; }
    *radix_log = 2; *s = S - 1 - 2 * step;
  } else {
    *radix_log = 2; *s = S - 2 - 2 * step;
  }
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
// inverse table from the forward one: Tinv_k[j] = omega^-j = -T_k[2^k - j]  (j > 0),  Tinv_k[0] = 1
B381_DI void ntt_inverse_twiddle(uint64_t idx, const fr_t* fwd, fr_t* inv_table) {
  uint32_t k = 63;
  while (!(((idx + 1) >> k) & 1)) k--;
  const uint64_t j = idx + 1 - (1ull << k);
  inv_table[idx] = j ? neg(fwd[(1ull << k) - 1 + ((1ull << k) - j)]) : fwd[idx];
}

// ---- pass planner (host) --------------------------------------------------------------------
// 1024 elements = 32 KB of shared memory per CTA, 256 threads (one radix-4 group per thread and step), 3 CTAs per SM.
// Measured on B200 against the 2048-element tile of round 1 (same threads and CTAs per SM, two groups per thread):
// 2^24 3.55 -> 3.45 ms, 2^22 0.852 -> 0.826, 2^20 0.250 -> 0.222 (now three passes), 256 x 2^16 2.27 -> 2.21
// (profiles/r02_ntt_variants.txt): the 96 KB of shared memory the smaller tiles leave go to the L1, which serves the
// twiddle reads.  128 threads x 6 CTAs on the same tile: 3.74 ms.
constexpr uint32_t kNttTileLog = 10;
struct ntt_pass_plan { uint32_t lo, S, g, x; };

// Splits the n stages into passes for a 2^tl-element tile.  The last pass (lo = 0) works on contiguous tiles and may
// take up to tl stages; earlier passes take at most tl - 2, so that a tile row keeps >= 4 adjacent elements (128 B
// contiguous) for coalescing.  Among the splits with the fewest passes, stage counts are balanced and EVEN where
// possible (an odd count costs one extra radix-2 step = one more trip through shared memory): 24 -> 8 + 8 + 8,
// 22 -> 6 + 8 + 8, 20 -> 9 + 11.  Fills `out` (capacity 8), highest stages first, and returns the number of passes.
// explicit split (stage counts, highest stages first; sum must be n): experiments and tests
inline int ntt_plan_from_list(uint32_t n, const uint32_t* S, uint32_t P, ntt_pass_plan* out, uint32_t tl = kNttTileLog) {
  uint32_t hi = n, sum = 0;
  for (uint32_t i = 0; i < P; i++) sum += S[i];
  if (sum != n || P < 1 || P > 8) return 0;
  for (uint32_t i = 0; i < P; i++) {
    if (S[i] < 1 || S[i] > tl || (i + 1 < P && S[i] >= tl)) return 0;
    if (i + 1 < P) out[i] = ntt_pass_plan{hi - S[i], S[i], tl - S[i], 0};
    else out[i] = ntt_pass_plan{0, S[i], 0, tl - S[i]};
    if (i + 1 < P && out[i].g > out[i].lo) return 0;
    hi -= S[i];
  }
  return (int)P;
}

inline int ntt_plan_passes(uint32_t n, ntt_pass_plan* out, uint32_t tl = kNttTileLog) {
  int np = 0;
  if (n <= tl) {
    out[np++] = ntt_pass_plan{0, n, 0, tl - n};
    return np;
  }
  const uint32_t up_max = tl - 2;
  uint32_t P = 2;
  while ((P - 1) * up_max + tl < n) P++;
  uint32_t S[8];
  auto cap = [&](uint32_t i) { return i + 1 == P ? tl : up_max; };
  // balanced start within the per-pass limits: every pass floor(n / P), remainder to the last passes, overflow of a
  // capped pass handed on to the next one
  uint32_t carry = 0;
  for (uint32_t i = 0; i < P; i++) {
    uint32_t want = n / P + ((i >= P - n % P) ? 1u : 0u) + carry;
    S[i] = want > cap(i) ? cap(i) : want;
    carry = want - S[i];
  }
  auto ok = [&](uint32_t i, uint32_t v) { return v >= 1 && v <= cap(i); };
  // then make counts even by moving single stages between an odd pass and the next odd pass, where the limits allow
  for (uint32_t i = 0; i + 1 < P; i++) {
    if ((S[i] & 1) == 0) continue;
    for (uint32_t j = i + 1; j < P; j++) {
      if (!(S[j] & 1)) continue;
      if (ok(i, S[i] - 1) && ok(j, S[j] + 1)) { S[i]--; S[j]++; }
      else if (ok(i, S[i] + 1) && ok(j, S[j] - 1)) { S[i]++; S[j]--; }
      break;
    }
  }
  uint32_t hi = n;
  for (uint32_t i = 0; i < P; i++) {
    if (i + 1 < P) out[np++] = ntt_pass_plan{hi - S[i], S[i], tl - S[i], 0};
    else out[np++] = ntt_pass_plan{0, S[i], 0, tl - S[i]};
    hi -= S[i];
  }
  return np;
}

// Column passes of the distributed (four-step) transform: the top `a` DIF stages of a 2^log_n transform on rank's
// column block (n_loc = log_n - log_gpus local index bits), split into passes of <= tl - 2 stages with
// g = min(tl - S, logL, lo) adjacent columns.  Fills `out` (capacity 8) with everything but the peer pointers.
inline int ntt_dist_passes(uint32_t log_n, uint32_t log_gpus, uint32_t rank, uint32_t a, const fr_t* tw,
                           ntt_pass_params* out, uint32_t tl = kNttTileLog) {
  const uint32_t lo_glob = log_n - a;           // global bit where the column index ends
  const uint32_t logL = lo_glob - log_gpus, n_loc = log_n - log_gpus;
  const uint32_t up_max = tl - 2;
  uint32_t rest = a, np = (a + up_max - 1) / up_max, hi = n_loc;
  for (uint32_t i = 0; i < np; i++) {
    const uint32_t S = (rest + (np - i) - 1) / (np - i);
    ntt_pass_params p;
    memset(&p, 0, sizeof(p));
    p.n = n_loc; p.lo = hi - S; p.S = S;
    p.g = tl - S;
    if (p.g > logL) p.g = logL;
    if (p.g > p.lo) p.g = p.lo;
    p.total = 1ull << n_loc;
    p.estride = 1; p.bstride = 1ull << n_loc;
    p.twiddles = tw;
    p.dist_shift = log_gpus; p.dist_logL = logL; p.dist_lo = lo_glob; p.dist_lbase = rank << logL;
    out[i] = p;
    hi -= S;
    rest -= S;
  }
  return (int)np;
}

}  // namespace b381
