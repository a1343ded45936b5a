// Affine bucket pre-reduction with batched inversion (G1 and G2) -- per-thread bodies.
//
// After the sort every bucket is a contiguous run of (point, sign) entries.  One LEVEL replaces
// each bucket of s points by ceil(s/2) points: neighbours (2i, 2i+1) are added in AFFINE
// coordinates, lambda = (y2-y1)/(x2-x1), the denominators inverted together with Montgomery's
// trick: 3 products per pair for the shared inversion + 2M + 1S for the addition =
// 6 Fq products per removed point, against 10 (8M+2S) for the XYZZ mixed addition the serial
// accumulate kernel spends -- and the work is spread by OUTPUT SLOT, not by bucket, so it is
// balanced whatever the scalar distribution.  After a few levels the buckets are short and the
// existing task/accumulate/finalize path finishes them.
//
// The reference has no counterpart: it accumulates in Jacobian coordinates with an add that also
// evaluates a doubling every time (bls12-381/src/curve/msm_kernels.cu:269-366,
// include/point.cuh:803-912).  The affine result of the MSM is representation independent, so
// bit-exactness of the result is unaffected.
//
// One thread owns B consecutive output slots.  Three kernels per level, no barrier anywhere:
//   k_msm_pair_fwd     walk the slots (slot -> input pair), denominators, running product; the exclusive
//                      prefixes and the thread's total go to global memory (slot-major, coalesced)
//   k_msm_invert_totals  Montgomery's trick once more over M totals per thread + ONE variable-time
//                      inversion per thread (field.cuh inv_vartime): ~3.4 products per total
//   k_msm_pair_bwd     walk backward: 1/d_k = inv * prefix_k, inv *= d_k, finish the affine addition
// (A first version inverted once per CTA behind __syncthreads; ncu showed the barrier as the top stall
// and the IMAD pipe 50 % busy, so the inversion moved into its own kernel.)
#pragma once
#include "curve.cuh"

namespace b381 {

// launch shape of the per-level kernels (msm_pair.cu) and slots per thread / max totals per inversion
constexpr int PR_TPB = 128;
template <class F> struct pair_batch { static constexpr int B = 32, M = 64; };
template <> struct pair_batch<fq2_t> { static constexpr int B = 16, M = 32; };

enum : uint32_t { PAIR_NONE = 0xFFFFFFFFu, PAIR_SINGLE = 0x80000000u };
enum : int { PK_COPY_P = 0, PK_COPY_Q = 1, PK_INF = 2, PK_ADD = 3, PK_DBL = 4 };

// 256-bit global loads (LDG.E.256, new on sm_100) for the point gathers: a 96-byte G1 point is 3 requests instead
// of 6, its x-coordinate 2 instead of 3.  The level-0 forward pass is bound by the RATE of divergent line requests
// (every lane of a load instruction hits its own 128-byte line; profiles/r01b_msm_levels_sweep.txt), not by bytes.
// Needs 32-byte alignment of the element (cudaMalloc'ed point arrays: 96- and 192-byte points keep it); anything
// else takes the plain path.  Read-only data within the kernel, so the non-coherent path is legal.
template <class T>
B381_DI T load_wide(const T* p) {
#if defined(__CUDA_ARCH__) && !defined(B381_NO_WIDE_LD)
  static_assert(sizeof(T) % 16 == 0, "16-byte granules");
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    T r;
    uint64_t* w = reinterpret_cast<uint64_t*>(&r);
    const char* a = reinterpret_cast<const char*>(p);
    constexpr int N32 = (int)(sizeof(T) / 32);
#pragma unroll
    for (int i = 0; i < N32; i++)
      asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
                   : "=l"(w[4 * i]), "=l"(w[4 * i + 1]), "=l"(w[4 * i + 2]), "=l"(w[4 * i + 3]) : "l"(a + 32 * i));
    if (sizeof(T) % 32)
      asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(w[4 * N32]), "=l"(w[4 * N32 + 1]) : "l"(a + 32 * N32));
    return r;
  }
#endif
  return *p;
}

// 128-bit loads of an element that is only 16-byte aligned (a coordinate inside a struct-of-arrays level buffer)
template <class T>
B381_DI T load_vec16(const T* p) {
#if defined(__CUDA_ARCH__) && !defined(B381_NO_WIDE_LD)
  static_assert(sizeof(T) % 16 == 0, "16-byte granules");
  T r;
  uint64_t* w = reinterpret_cast<uint64_t*>(&r);
  const char* a = reinterpret_cast<const char*>(p);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++)
    asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(w[2 * i]), "=l"(w[2 * i + 1]) : "l"(a + 16 * i));
  return r;
#else
  return *p;
#endif
}

template <class T>
B381_DI void store_vec16(T* p, const T& v) {
#if defined(__CUDA_ARCH__) && !defined(B381_NO_WIDE_LD)
  const uint64_t* w = reinterpret_cast<const uint64_t*>(&v);
  char* a = reinterpret_cast<char*>(p);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++)
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(a + 16 * i), "l"(w[2 * i]), "l"(w[2 * i + 1]) : "memory");
#else
  *p = v;
#endif
}

// What a level reads.  Level 0 gathers the caller's bases (array of 96/192-byte structs) through the sorted entries;
// every later level reads the previous level's output, which is kept as TWO arrays (x[], y[]): the forward pass needs
// the x-coordinates only, and as structs every 48-byte x drags its y through DRAM as well (levels >= 1 of a 2^24-point
// MSM: 6.7 ms of forward passes at 66 % DRAM).
template <class F> struct level_pts {
  const affine_t<F>* aos;   // level 0
  const F* x;               // levels >= 1
  const F* y;
};
template <class F> B381_HD level_pts<F> level_from_bases(const affine_t<F>* b) { return level_pts<F>{b, nullptr, nullptr}; }
template <class F> B381_HD level_pts<F> level_from_xy(const F* x, const F* y) { return level_pts<F>{nullptr, x, y}; }

// point at sorted position `pos` of the current level; level 0 gathers through the sorted
// (index, sign) entries and applies the sign
template <class F, bool L0>
B381_DI affine_t<F> pair_load(uint32_t pos, const uint32_t* svals, const level_pts<F>& pts) {
  if (L0) {
    uint32_t v = svals[pos];
    affine_t<F> p = load_wide(pts.aos + (v >> 1));
    if ((v & 1) && !is_inf(p)) p.y = neg(p.y);
    return p;
  }
  return affine_t<F>{load_vec16(pts.x + pos), load_vec16(pts.y + pos)};
}

// what P + Q needs, and the denominator that goes into the batched inversion
template <class F>
B381_DI int pair_classify(const affine_t<F>& p, const affine_t<F>& q, F& denom) {
  denom = one<F>();
  if (is_inf(q)) return PK_COPY_P;
  if (is_inf(p)) return PK_COPY_Q;
  F dx = sub(q.x, p.x);
  if (is_zero(dx)) {
    if (!eq(p.y, q.y)) return PK_INF;       // P + (-P)
    denom = dbl(p.y);                        // y != 0 on a curve without 2-torsion
    return PK_DBL;
  }
  denom = dx;
  return PK_ADD;
}

template <class F>
B381_DI affine_t<F> pair_finish(int kind, const affine_t<F>& p, const affine_t<F>& q, const F& dinv) {
  if (kind == PK_COPY_P) return p;
  if (kind == PK_COPY_Q) return q;
  if (kind == PK_INF) return affine_t<F>{zero<F>(), zero<F>()};
  F num;
  if (kind == PK_ADD) num = sub(q.y, p.y);
  else { F x2 = sqr(p.x); num = add(dbl(x2), x2); }
  F lam = mul(num, dinv);
  affine_t<F> r;
  r.x = sub(sub(sqr(lam), p.x), q.x);
  r.y = sub(mul(lam, sub(p.x, r.x)), p.y);
  return r;
}

// ceil(size/2) per bucket; the exclusive scan of these is the next level's offsets
B381_DI void msm_half_counts_body(uint32_t b, const uint32_t* offsets, uint32_t nbuckets, uint32_t Bs, uint32_t* counts) {
  if (b > nbuckets) return;
  // trash slots (b % Bs == Bs - 1) produce nothing, so they are empty from level 1 on
  counts[b] = (b < nbuckets && b % Bs != Bs - 1) ? (offsets[b + 1] - offsets[b] + 1) / 2 : 0u;
}

// Chunk-major level 0 (msm_core.cuh): run r = (b * nchunks + chunk) * S + slot (S slots and nchunks chunks per MSM).
// Its sums are WRITTEN bucket-major, i.e. in the order (b, slot, chunk): position perm(r) = (b * S + slot) * nchunks +
// chunk of the second count array, whose exclusive scan gives every run's first output position -- and, at chunk 0,
// the level-1 offsets of the bucket slots.
B381_HD uint32_t msm_run_perm(uint32_t r, uint32_t set_slots, uint32_t nchunks) {
  const uint32_t g = r / set_slots, slot = r - g * set_slots;
  const uint32_t b = g / nchunks, chunk = g - b * nchunks;
  return (b * set_slots + slot) * nchunks + chunk;
}
B381_DI void msm_half_counts_runs_body(uint32_t r, const uint32_t* run_off, uint32_t nbuckets, uint32_t set_slots,
                                       uint32_t nchunks, uint32_t Bs, uint32_t* counts, uint32_t* counts_bucket_major,
                                       uint32_t r_end = 0xFFFFFFFFu) {
  const uint32_t nruns = nbuckets * nchunks;
  if (r_end > nruns) r_end = nruns;
  if (r > r_end) return;                        // r_end < nruns: one piece of runs at a time (streamed level 0)
  if (r == r_end) {                             // sentinel: the scan leaves the slot total of [.., r_end) here
    counts[r] = 0;
    if (r == nruns) counts_bucket_major[r] = 0;
    return;
  }
  const uint32_t h = (r % Bs != Bs - 1) ? (run_off[r + 1] - run_off[r] + 1) / 2 : 0u;   // set_slots is a multiple of Bs
  counts[r] = h;
  counts_bucket_major[msm_run_perm(r, set_slots, nchunks)] = h;
}

// Streamed level 0 (msm_impl.cuh): the forward pass runs once per piece of runs.  Does thread t (slots [t*B, t*B + B))
// run in the piece that brought the slot total from lo to hi?  It does when the piece completes its B slots; the last
// piece also takes the thread with the ragged end.
B381_HD bool pair_piece_owns(uint32_t t, uint32_t B, uint32_t lo, uint32_t hi, bool final) {
  const uint64_t first = (uint64_t)t * B;
  if (first + B <= lo) return false;            // done with an earlier piece
  return final ? first < hi : first + B <= hi;
}

// slot -> input pair.  src[k*stride] = position of the pair's first point (| PAIR_SINGLE when the bucket's
// odd last point is just carried over), PAIR_NONE past the end.
// dst_base != nullptr (chunk-major level 0): "buckets" are runs, and dst[k*stride] = where the slot's sum goes:
// dst_base[perm(run)] + (slot - first slot of the run).
template <int B>
B381_DI void pair_walk(uint32_t slot0, uint32_t n_out, const uint32_t* in_off, const uint32_t* out_off,
                       uint32_t nbuckets, uint32_t* src, size_t stride, const uint32_t* dst_base = nullptr,
                       uint32_t dst_set_slots = 0, uint32_t dst_nchunks = 0, uint32_t* dst = nullptr) {
  uint32_t b = 0, ob = 0, oe = 0, ib = 0, ie = 0, db = 0;
  if (slot0 < n_out) {
    // last bucket whose first output slot is <= slot0 (empty buckets share a start with their
    // successor, so "last" is the one that really owns the slot)
    uint32_t lo = 0, hi = nbuckets;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (out_off[mid] <= slot0) lo = mid; else hi = mid;
    }
    b = lo;
    ob = out_off[b]; oe = out_off[b + 1];
    ib = in_off[b]; ie = in_off[b + 1];
    if (dst_base) db = dst_base[msm_run_perm(b, dst_set_slots, dst_nchunks)];
  }
#ifndef B381_HOST_TEST
#pragma unroll 1
#endif
  for (int k = 0; k < B; k++) {
    uint32_t j = slot0 + (uint32_t)k;
    uint32_t s = PAIR_NONE;
    if (j < n_out) {
      if (j >= oe) {
        do { b++; ob = oe; oe = out_off[b + 1]; } while (j >= oe);
        ib = in_off[b]; ie = in_off[b + 1];
        if (dst_base) db = dst_base[msm_run_perm(b, dst_set_slots, dst_nchunks)];
      }
      s = ib + 2u * (j - ob);
      if (s + 1 >= ie) s |= PAIR_SINGLE;
      if (dst_base) dst[(size_t)k * stride] = db + (j - ob);
    }
    if (src) src[(size_t)k * stride] = s;        // nullptr: destinations only (k_msm_pair_dst)
  }
}

// x-coordinates alone, one per 64-byte (Fq) / 128-byte (Fq2) aligned record: the level-0 forward pass gathers
// them at random and needs nothing else, so each gather is one aligned line instead of a 96-byte struct that
// straddles two (k_pack_x in msm_pair.cu builds the array per call: 0.4 ms at 2^24).
template <class F> struct alignas(sizeof(F) <= 64 ? 64 : 128) xrec_t { F x; };

template <class F, bool L0>
B381_DI F pair_load_x(uint32_t pos, const uint32_t* svals, const level_pts<F>& pts, const xrec_t<F>* xs = nullptr) {
  if (L0) {
    uint32_t i = svals[pos] >> 1;
    return xs ? xs[i].x : load_wide(&pts.aos[i].x);
  }
  return load_vec16(pts.x + pos);
}

// Forward: pre[k*pstride] = product of the denominators before k; returns the product of all B.
// Software-pipelined three slots deep, because a slot is a chain of three dependent loads (slot word -> sorted entries
// -> x-coordinates, the last one a random gather at level 0) followed by one 430-instruction product, and ptxas keeps
// that chain inside the iteration: round 1's loop ran at 42 % of the multiplier pipe with long_scoreboard as top stall.
// Here iteration k issues the slot word of k+3, the entries of k+2 and the x-gathers of k+1 BEFORE multiplying slot k.
template <class F, int B, bool L0>
B381_DI F pair_phase1(const uint32_t* src, size_t sstride, const uint32_t* svals, const level_pts<F>& pts, F* pre,
                      size_t pstride, const xrec_t<F>* xs = nullptr) {
  F acc = one<F>();
  auto slot_word = [&](int k) -> uint32_t { return k < B ? src[(size_t)k * sstride] : (uint32_t)PAIR_NONE; };
  // PAIR_NONE has the PAIR_SINGLE bit set: neither kind of slot has a denominator
  auto entries = [&](uint32_t s, uint32_t& i1, uint32_t& i2) {
    if (s & PAIR_SINGLE) return;
    if (L0) { i1 = svals[s] >> 1; i2 = svals[s + 1] >> 1; }
    else { i1 = s; i2 = s + 1; }
  };
  auto gather = [&](uint32_t s, uint32_t i1, uint32_t i2, F& x1, F& x2) {
    if (s & PAIR_SINGLE) return;
    if (L0 && xs) { x1 = xs[i1].x; x2 = xs[i2].x; }
    else if (L0) { x1 = load_wide(&pts.aos[i1].x); x2 = load_wide(&pts.aos[i2].x); }
    else { x1 = load_vec16(pts.x + i1); x2 = load_vec16(pts.x + i2); }
  };
  uint32_t s0 = slot_word(0), s1 = slot_word(1), s2 = slot_word(2);
  uint32_t a1 = 0, a2 = 0, b1 = 0, b2 = 0;            // entries of slot k (a) and k+1 (b)
  F x1 = zero<F>(), x2 = zero<F>(), y1 = zero<F>(), y2 = zero<F>();   // x-coordinates of slot k (x) and k+1 (y)
  entries(s0, a1, a2);
  gather(s0, a1, a2, x1, x2);
  entries(s1, b1, b2);
#ifndef B381_HOST_TEST
#pragma unroll 2
#endif
  for (int k = 0; k < B; k++) {
    const uint32_t s3 = slot_word(k + 3);
    uint32_t c1 = 0, c2 = 0;
    entries(s2, c1, c2);                  // slot k+2
    gather(s1, b1, b2, y1, y2);           // slot k+1
    pre[(size_t)k * pstride] = acc;
    if (!(s0 & PAIR_SINGLE)) {
      F d = sub(x2, x1);
      if (is_zero(x1) || is_zero(x2) || is_zero(d)) {   // infinity operand, doubling or cancellation: rare
        affine_t<F> p = pair_load<F, L0>(s0, svals, pts);
        affine_t<F> q = pair_load<F, L0>(s0 + 1, svals, pts);
        pair_classify(p, q, d);
      }
      acc = mul(acc, d);
    }
    s0 = s1; s1 = s2; s2 = s3;
    a1 = b1; a2 = b2; b1 = c1; b2 = c2;
    x1 = y1; x2 = y2;
  }
  return acc;
}

// Backward: `inv_total` is the inverse of what phase1 returned; outx / outy point at the thread's first slot.
// dst == nullptr: slot k's sum goes to outx/outy[k] (the caller passes the thread's first slot); else to
// outx/outy[dst[k*sstride]] (chunk-major level 0: pair_walk).
template <class F, int B, bool L0>
B381_DI void pair_phase2(F inv_total, const uint32_t* src, size_t sstride, const uint32_t* svals,
                         const level_pts<F>& pts, const F* pre, size_t pstride, F* outx, F* outy,
                         const uint32_t* dst = nullptr) {
  uint32_t s_next = src[(size_t)(B - 1) * sstride];
#pragma unroll 1
  for (int k = B - 1; k >= 0; k--) {
    uint32_t s = s_next;
    if (k) s_next = src[(size_t)(k - 1) * sstride];   // one hop of the src -> svals -> point chain ahead
    if (s == PAIR_NONE) continue;
    uint32_t p0 = s & ~PAIR_SINGLE;
    const size_t o = dst ? (size_t)dst[(size_t)k * sstride] : (size_t)k;
    affine_t<F> p = pair_load<F, L0>(p0, svals, pts);
    if (s & PAIR_SINGLE) { store_vec16(outx + o, p.x); store_vec16(outy + o, p.y); continue; }
    affine_t<F> q = pair_load<F, L0>(p0 + 1, svals, pts);
    F d;
    int kind = pair_classify(p, q, d);
    F dinv = mul(inv_total, pre[(size_t)k * pstride]);
    if (k) inv_total = mul(inv_total, d);
    const affine_t<F> r = pair_finish(kind, p, q, dinv);
    store_vec16(outx + o, r.x);
    store_vec16(outy + o, r.y);
  }
}

// In-place inversion of `count` field elements, element j of thread t at v[j*nthreads + t], m <= MAXM
// elements per thread: prefix products, one inversion, unwind.
template <class F, int MAXM>
B381_DI void batch_invert_body(uint32_t t, uint32_t nthreads, uint32_t count, uint32_t m, F* v) {
  if (t >= nthreads) return;
  F pf[MAXM];
  F acc = one<F>();
#pragma unroll 1
  for (uint32_t j = 0; j < m; j++) {
    size_t idx = (size_t)j * nthreads + t;
    pf[j] = acc;
    if (idx < count) acc = mul(acc, v[idx]);
  }
  F inv_acc = inv_vartime(acc);
#pragma unroll 1
  for (int j = (int)m - 1; j >= 0; j--) {
    size_t idx = (size_t)j * nthreads + t;
    if (idx >= count) continue;
    F x = v[idx];
    v[idx] = mul(inv_acc, pf[j]);
    inv_acc = mul(inv_acc, x);
  }
}

// elements per thread of the batched inversion: as many as keep ~2 waves of threads busy, at most MAXM
B381_HD uint32_t batch_invert_m(uint32_t count, uint32_t maxm) {
  uint32_t m = count / (148u * 256u);
  if (m < 4) m = 4;
  if (m > maxm) m = maxm;
  return m;
}

}  // namespace b381
