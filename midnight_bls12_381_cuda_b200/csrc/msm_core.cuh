// Per-thread bodies of the Pippenger MSM pipeline (G1 and G2), written so that each CUDA
// kernel in msm_impl.cuh is `body(global_thread_id, ...)`.  The same bodies are driven by a serial
// loop in tests/host/msm_host_sim.cpp (CPU-only CI), which is how the pipeline logic is
// checked without a GPU.
//
// Pipeline (ours; the reference's is bls12-381/src/curve/msm_kernels.cu:603-903):
//   1 digits      scalar -> W signed c-bit digits -> (key = w*(B+1) + |d|-1, val = idx<<1 | sign);
//                 zero digits go to the window's own trash bucket, key w*(B+1) + B
//                                                           [ref: compute_bucket_indices_kernel :69-143]
//   2 sort        keys are written window-major, so the data is already grouped by window: ONE radix sort
//                 PER WINDOW over the c bits of the in-window key (2 onesweep passes at c = 16), or, for small
//                 n / folded windows, one sort over ceil(log2(W*(B+1))) bits   [ref sorts all 32 bits, :768-778]
//   3 offsets     bucket boundaries straight from the sorted keys (no histogram, no atomics)
//                                                           [ref: histogram + scan, :224-256,:748-758]
//   3b pre-reduce affine pairwise levels with CTA-wide batched inversion (msm_batch.cuh): each level
//                 halves every bucket at 6 Fq products per removed point      [no reference counterpart]
//   4 tasks       every bucket is cut into ceil(size/K) equal tasks so one thread never owns more
//                 than K insertions, whatever the scalar distribution
//                                                           [ref: 1 thread or 8/16 threads per bucket]
//   5 accumulate  one thread per task: XYZZ mixed additions (8M+2S)     <- the hot loop
//   6 finalize    bucket = sum of its task partials
//   7 seg-reduce  segments of L buckets: T_s + (s*L)*A_s  (running-sum trick + short double-and-add)
//                                                           [ref: 1 CTA per window, :376-513]
//   8 tree        pairwise sums down to one point per window
//   9 combine     Horner over windows, to-affine, ICICLE standard-form (x,y,1) result
//                                                           [ref: :529-596 + icicle_curve_api.cu:134-229]
#pragma once
#include "curve.cuh"
#include "msm_batch.cuh"

namespace b381 {

struct msm_shape {
  uint32_t n;        // points in this MSM
  uint32_t c;        // window bits
  uint32_t W;        // number of c-bit windows of the scalar
  uint32_t B;        // buckets per window = 2^(c-1)
  uint32_t Wf;       // bucket sets after folding with precomputed bases: ceil(W / precompute_factor)
  uint32_t Bs;       // bucket slots per set = B + 1: slot B of every set is its trash bucket (zero digits)
  uint32_t nbuckets; // Wf*Bs bucket slots, trash slots included
  uint32_t f;        // precompute_factor (>= 1): point i has f stored multiples, at bases[i*f + k]
};

// With precompute_factor f the caller supplies f*n bases, point i's multiples 2^(k*Wf*c) * P_i (k < f) stored
// INTERLEAVED at bases[i*f + k] -- upstream ICICLE's layout, which lets an MSM over the first n' < n points use a
// prefix of the same buffer (core/msm.rs:654-661 passes the full buffer whatever scalars.len() is).  Window
// w = k*Wf + w' is inserted into bucket set w' using multiple k -- all windows of one residue class share buckets,
// so only Wf bucket sets are reduced and combined.
B381_HD msm_shape make_msm_shape(uint32_t n, uint32_t c, uint32_t bits, uint32_t factor) {
  msm_shape sh;
  sh.n = n;
  sh.c = c;
  sh.W = (bits + 1 + c - 1) / c;
  sh.B = 1u << (c - 1);
  if (factor < 1) factor = 1;
  sh.Wf = (sh.W + factor - 1) / factor;
  sh.Bs = sh.B + 1;
  sh.nbuckets = sh.Wf * sh.Bs;
  sh.f = factor;
  return sh;
}

B381_HD uint32_t ceil_div_u32(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// A bucket is cut into at most this many tasks, however large it is (all scalars equal puts every point
// of a window into ONE bucket): its partials are then summed by one warp (32 per lane + a shuffle tree,
// msm_tail.cu) instead of a serial chain of thousands of additions.
constexpr uint32_t kMaxTasksPerBucket = 1024;
// buckets with more partials than this are left to the warp-per-bucket finalize kernel
constexpr uint32_t kFinalizeSerialMax = 8;
// entries of bucket slot b that count: the trash slot of every set (b % Bs == Bs - 1) holds the zero digits
B381_HD uint32_t msm_bucket_size(const uint32_t* offsets, uint32_t b, uint32_t Bs) {
  return (b % Bs == Bs - 1) ? 0u : offsets[b + 1] - offsets[b];
}
B381_HD uint32_t msm_tasks_of(uint32_t size, uint32_t K) {
  uint32_t nt = ceil_div_u32(size, K);
  return nt > kMaxTasksPerBucket ? kMaxTasksPerBucket : nt;
}

// ---------------------------------------------------------------- 1 digits
// Signed-digit recoding with the same digit set as the reference
// (msm_kernels.cu:96-130): d in [-(2^(c-1)-1), 2^(c-1)], d > 2^(c-1) => d -= 2^c, carry.
// keys/vals are window-major ([w*n + i]) so stores coalesce.
// local_keys: emit the in-window key only (0..B, B = trash) -- the per-window sort path (needs Wf == W).
B381_DI void msm_digits_body(uint32_t i, const fr_t* scalars, bool scalars_mont, const msm_shape sh,
                             uint32_t* keys, uint32_t* vals, bool local_keys = false) {
  if (i >= sh.n) return;
  fr_t s = scalars[i];
  if (scalars_mont) s = from_mont(s);
  const uint32_t mask = (1u << sh.c) - 1u;
  uint32_t carry = 0;
  for (uint32_t w = 0; w < sh.W; w++) {
    uint32_t bit = w * sh.c;
    uint32_t limb = bit >> 6, off = bit & 63;
    uint32_t d = 0;
    if (limb < 4) {
      uint64_t lo = s.l[limb] >> off;
      if (off + sh.c > 64 && limb + 1 < 4) lo |= s.l[limb + 1] << (64 - off);
      d = (uint32_t)lo & mask;
    }
    d += carry;
    carry = 0;
    uint32_t sign = 0;
    if (d > sh.B) {          // B = 2^(c-1)
      d = (1u << sh.c) - d;
      sign = 1;
      carry = 1;
    }
    uint32_t blk = w / sh.Wf, wf = w - blk * sh.Wf;
    uint32_t key = (local_keys ? 0u : wf * sh.Bs) + (d ? d - 1 : sh.B);
    keys[(size_t)w * sh.n + i] = key;
    vals[(size_t)w * sh.n + i] = ((i * sh.f + blk) << 1) | sign;
  }
  // W*c >= 256 > bit length of any canonical scalar, so the last carry is always 0.
}

// ---------------------------------------------------------------- 3 offsets
// sorted keys -> offsets[0..nbuckets]; offsets[b] = first position with key >= b; offsets[nbuckets] = total.
// n_local != 0: the keys are in-window keys, sorted per window slice of n_local entries; the bucket slot of
// position j is (j / n_local) * Bs + key, which is monotone over the concatenated slices.
// position j of the sorted keys, whose neighbours j-1 / j belong to windows w_prev / w_cur (per-window sort: keys
// are in-window, the bucket-set base w * Bs is added here; 0 otherwise)
B381_DI void msm_offsets_core(size_t j, uint32_t w_prev, uint32_t w_cur, const uint32_t* sorted_keys, size_t total,
                              uint32_t nbuckets, uint32_t* offsets, uint32_t Bs) {
  if (j > total) return;
  uint32_t kp = 0, kc = 0;
  if (j > 0) kp = sorted_keys[j - 1] + w_prev * Bs;
  if (j < total) kc = sorted_keys[j] + w_cur * Bs;
  uint32_t prev = (j == 0) ? 0u : kp + 1u;          // first key not yet started
  uint32_t cur = (j == total) ? nbuckets + 1u : kc + 1u; // one past this key
  // every bucket id in [prev, cur) starts at position j
  for (uint32_t b = prev; b < cur && b <= nbuckets; b++) offsets[b] = (uint32_t)j;
}
B381_DI void msm_offsets_body(size_t j, const uint32_t* sorted_keys, size_t total, uint32_t nbuckets,
                              uint32_t* offsets, uint32_t n_local = 0, uint32_t Bs = 0) {
  if (j > total) return;
  const uint32_t wp = (n_local && j > 0) ? (uint32_t)((j - 1) / n_local) : 0u;
  const uint32_t wc = n_local ? (uint32_t)(j / n_local) : 0u;
  msm_offsets_core(j, wp, wc, sorted_keys, total, nbuckets, offsets, n_local ? Bs : 0u);
}

// ---------------------------------------------------------------- 4 tasks
B381_DI void msm_task_count_body(uint32_t b, const uint32_t* offsets, uint32_t nbuckets, uint32_t Bs, uint32_t K,
                                 uint32_t* counts) {
  if (b >= nbuckets) return;
  uint32_t sz = msm_bucket_size(offsets, b, Bs);
  counts[b] = msm_tasks_of(sz, K);
}

B381_DI void msm_build_tasks_body(uint32_t b, const uint32_t* offsets, const uint32_t* task_start,
                                  uint32_t nbuckets, uint32_t Bs, uint32_t K, uint2* tasks) {
  if (b >= nbuckets) return;
  uint32_t beg = offsets[b], sz = msm_bucket_size(offsets, b, Bs);
  uint32_t nt = msm_tasks_of(sz, K);
  uint32_t t0 = task_start[b];
  // equal split: first (sz % nt) tasks get one extra element
  uint32_t base = nt ? sz / nt : 0, rem = nt ? sz % nt : 0, pos = beg;
  for (uint32_t t = 0; t < nt; t++) {
    uint32_t len = base + (t < rem ? 1u : 0u);
    uint2 tk;
    tk.x = pos;
    tk.y = pos + len;
    tasks[t0 + t] = tk;
    pos += len;
  }
}

// sort key of task t for the longest-first visiting order: K - len (valid), K + 1 (unused slot)
B381_DI void msm_task_key_body(uint32_t t, uint32_t max_tasks, uint32_t ntasks, const uint2* tasks, uint32_t K,
                               uint32_t* keys, uint32_t* ids) {
  if (t >= max_tasks) return;
  uint32_t key = K + 1;
  if (t < ntasks) {
    uint2 tk = tasks[t];
    uint32_t len = tk.y - tk.x;
    key = len < K ? K - len : 0u;        // capped buckets have tasks longer than K: visit them first
  }
  keys[t] = key;
  ids[t] = t;
}

// ---------------------------------------------------------------- 5 accumulate (hot)
template <class F>
B381_DI void msm_accumulate_body(uint32_t t, uint32_t ntasks, const uint2* tasks, const uint32_t* sorted_vals,
                                 const affine_t<F>* bases, xyzz_t<F>* partial, const uint32_t* order = nullptr) {
  if (order) t = order[t];              // tasks visited longest-first so a warp's 32 tasks have similar lengths
  if (t >= ntasks) return;
  uint2 tk = tasks[t];
  xyzz_t<F> acc = xyzz_identity<F>();
  for (uint32_t j = tk.x; j < tk.y; j++) {
    uint32_t v = sorted_vals ? sorted_vals[j] : (j << 1);   // nullptr: `bases` is a pre-reduced level (msm_batch.cuh)
    affine_t<F> p = bases[v >> 1];
    if (is_inf(p)) continue;            // (0,0) bases are legal and contribute nothing
    if (v & 1) p.y = neg(p.y);
    xyzz_madd(acc, p);
  }
  partial[t] = acc;
}

// ---------------------------------------------------------------- 6 finalize
template <class F>
B381_DI void msm_finalize_body(uint32_t b, uint32_t nbuckets, const uint32_t* task_start,
                               const uint32_t* counts, const xyzz_t<F>* partial, xyzz_t<F>* buckets,
                               uint32_t serial_max = 0xFFFFFFFFu) {
  if (b >= nbuckets) return;
  uint32_t t0 = task_start[b], nt = counts[b];
  if (nt > serial_max) return;            // a warp does this bucket (k_msm_finalize_heavy)
  xyzz_t<F> acc = xyzz_identity<F>();
  if (nt) acc = partial[t0];
  for (uint32_t t = 1; t < nt; t++) xyzz_add(acc, partial[t0 + t]);
  buckets[b] = acc;
}

// ---------------------------------------------------------------- 7 segment reduce
// window w, segment s covers buckets j = s*L .. s*L+L-1 (weights j+1).
//   sum_j (j+1) B_j = T_s + (s*L) * A_s,   A_s = sum B_j,  T_s = sum (j-s*L+1) B_j
template <class F>
B381_DI xyzz_t<F> xyzz_mul_small(const xyzz_t<F>& p, uint32_t k) {
  xyzz_t<F> r = xyzz_identity<F>();
  if (k == 0 || is_inf(p)) return r;
  int top = 31;
  while (!((k >> top) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    r = xyzz_dbl(r);
    if ((k >> i) & 1) xyzz_add(r, p);
  }
  return r;
}

template <class F>
B381_DI void msm_segment_body(uint32_t gid, uint32_t W, uint32_t B, uint32_t L, const xyzz_t<F>* buckets,
                              xyzz_t<F>* seg_out) {
  uint32_t segs = B / L;
  if (gid >= W * segs) return;
  uint32_t w = gid / segs, s = gid % segs;
  const xyzz_t<F>* bk = buckets + (size_t)w * (B + 1) + (size_t)s * L;     // B + 1 slots per set (trash last)
  xyzz_t<F> run = xyzz_identity<F>();
  xyzz_t<F> tri = xyzz_identity<F>();
  for (int j = (int)L - 1; j >= 0; j--) {
    xyzz_add(run, bk[j]);
    xyzz_add(tri, run);
  }
  xyzz_t<F> sh = xyzz_mul_small(run, s * L);
  xyzz_add(tri, sh);
  seg_out[gid] = tri;
}

// ---------------------------------------------------------------- 8 tree
// in-place halving over groups: a[g*stride + t] += a[g*stride + t + half]
template <class F>
B381_DI void msm_tree_body(uint32_t gid, uint32_t groups, uint32_t stride, uint32_t half, xyzz_t<F>* a) {
  if (gid >= groups * half) return;
  uint32_t g = gid / half, t = gid % half;
  xyzz_t<F> x = a[(size_t)g * stride + t];
  xyzz_add(x, a[(size_t)g * stride + t + half]);
  a[(size_t)g * stride + t] = x;
}

// ---------------------------------------------------------------- 9 combine
// Horner over the W window sums (stride apart), result left in XYZZ.
template <class F>
B381_DI xyzz_t<F> msm_combine(const xyzz_t<F>* wsum, uint32_t stride, uint32_t W, uint32_t c) {
  xyzz_t<F> r = xyzz_identity<F>();
  for (int w = (int)W - 1; w >= 0; w--) {
    for (uint32_t k = 0; k < c; k++) r = xyzz_dbl(r);
    xyzz_add(r, wsum[(size_t)w * stride]);
  }
  return r;
}

// ICICLE result convention (icicle_curve_api.cu:134-179): (x, y, 1) in STANDARD form,
// identity = (0, 1, 0).  With `mont` the coordinates stay Montgomery and identity = (0, R, 0),
// which is the reference's Jacobian convention (point.cuh:469-486) for its extern "C" test API.
template <class F>
B381_DI jacobian_t<F> msm_result_encode(const xyzz_t<F>& r, bool mont) {
  affine_t<F> a = xyzz_to_affine(r);
  jacobian_t<F> o;
  if (is_inf(r)) {
    o.x = zero<F>();
    o.y = one<F>();
    o.z = zero<F>();
  } else {
    o.x = a.x;
    o.y = a.y;
    o.z = one<F>();
  }
  if (!mont) {
    o.x = from_mont(o.x);
    o.y = from_mont(o.y);
    o.z = from_mont(o.z);
  }
  return o;
}

}  // namespace b381
