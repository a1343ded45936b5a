// Per-thread bodies of the Pippenger MSM pipeline (G1 and G2), written so that each CUDA
// kernel in msm_impl.cuh is `body(global_thread_id, ...)`.  The same bodies are driven by a serial
// loop in tests/host/msm_host_sim.cpp (CPU-only CI), which is how the pipeline logic is
// checked without a GPU.
//
// Pipeline (ours; the reference's is bls12-381/src/curve/msm_kernels.cu:603-903):
//   1 histogram   scalar -> W signed c-bit digits -> bucket slot key = w*(B+1) + |d|-1 (zero digits go to the
//                 window's own trash slot w*(B+1) + B); hist[key]++   [ref: compute_bucket_indices_kernel :69-143]
//   2 scan        bucket boundaries = exclusive scan of the histogram (own kernels, msm_sort.cu)
//                                                           [ref: histogram of SORTED keys + scan, :224-256,:748-758]
//   3 scatter     digits recomputed; vals[cursor[key]++] = idx<<1 | sign: a counting sort, no key array, no
//                 library sort                              [ref: CUB radix sort of all 32 key bits, :768-778]
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
  uint32_t n;        // points per MSM
  uint32_t c;        // window bits
  uint32_t W;        // number of c-bit windows of the scalar
  uint32_t B;        // buckets per window = 2^(c-1)
  uint32_t Wf;       // bucket sets per MSM after folding with precomputed bases: ceil(W / precompute_factor)
  uint32_t Bs;       // bucket slots per set = B + 1: slot B of every set is its trash bucket (zero digits)
  uint32_t nbuckets; // batch*Wf*Bs bucket slots, trash slots included
  uint32_t f;        // precompute_factor (>= 1): point i has f stored multiples, at bases[i*f + k]
  uint32_t batch;    // MSMs folded into this pipeline run: MSM b owns bucket sets [b*Wf, (b+1)*Wf)
  uint32_t shared;   // 1: every MSM of the batch reads bases[0 .. n*f); 0: MSM b reads bases[b*n*f ..)
  uint32_t chunk_log, nchunks;   // grouping granularity: scalar i of an MSM belongs to its chunk i >> chunk_log; chunks PER MSM
};

// CHUNK-MAJOR GROUPING.  The level-0 gathers of the affine pre-reduction touch the bases at random: 2^28 gathers over
// 1.6 GB at 2^24 points, DRAM-bound at 128 bytes per 48-byte x-coordinate.  So the entries are grouped by
// (MSM of the batch, chunk of 2^chunk_log consecutive scalars of that MSM, bucket slot of that MSM) instead of by
// bucket slot alone: "run" r = (b * nchunks + chunk) * S + slot, S = Wf * Bs slots per MSM, nchunks chunks per MSM.
// Level 0 pairs entries inside a run and visits the runs chunk by chunk, so at any moment the whole GPU gathers from
// ONE 100 MB range of the bases, which the 126 MB L2 holds; it writes its sums bucket-major (a run's sums land behind
// those of the same bucket's earlier chunks), so from level 1 on nothing knows about chunks.
// A chunk's entries occupy a fixed region of the sorted array (2^chunk_log * W entries), so a chunk can be
// histogrammed, scanned and scattered on its own -- while later chunks of host-resident scalars are still crossing
// PCIe (msm_sort.cu).  nchunks == 1 is the plain bucket-major grouping.
B381_HD uint32_t msm_runs(const msm_shape& sh) { return sh.nchunks * sh.nbuckets; }
B381_HD uint32_t msm_set_slots(const msm_shape& sh) { return sh.Wf * sh.Bs; }          // S: bucket slots of ONE MSM
B381_HD void msm_shape_set_chunks(msm_shape& sh, uint32_t chunk_log) {
  sh.chunk_log = chunk_log;
  sh.nchunks = chunk_log >= 32 ? 1u : (uint32_t)(((uint64_t)sh.n + (1ull << chunk_log) - 1) >> chunk_log);
  if (sh.nchunks <= 1) { sh.nchunks = 1; sh.chunk_log = 31; }
}
// run of scalar i of MSM b whose digit maps to bucket slot `key` (key = b * S + slot, msm_digit_at)
B381_HD size_t msm_run_of(const msm_shape& sh, uint32_t b, uint32_t i, uint32_t key) {
  const uint32_t S = sh.Wf * sh.Bs;
  return ((size_t)b * sh.nchunks + (i >> sh.chunk_log)) * S + (key - b * S);
}

// With precompute_factor f the caller supplies f*n bases, point i's multiples 2^(k*Wf*c) * P_i (k < f) stored
// INTERLEAVED at bases[i*f + k] -- upstream ICICLE's layout, which lets an MSM over the first n' < n points use a
// prefix of the same buffer (core/msm.rs:654-661 passes the full buffer whatever scalars.len() is).  Window
// w = k*Wf + w' is inserted into bucket set w' using multiple k -- all windows of one residue class share buckets,
// so only Wf bucket sets are reduced and combined.
// A BATCH of MSMs (MSMConfig.batch_size, core/msm.rs:1179-1295) is one pipeline run over batch*Wf bucket sets: the
// grouping, the affine levels, the task/accumulate/finalize stages and the bucket reduction never look at which MSM a
// bucket slot belongs to, so the latency-bound tail (~3 ms) is paid once per batch instead of once per MSM, and small
// MSMs fill the GPU together.  (The reference loops over the batch on the host, icicle_curve_api.cu:243-407.)
B381_HD msm_shape make_msm_shape(uint32_t n, uint32_t c, uint32_t bits, uint32_t factor, uint32_t batch = 1,
                                 bool shared = true) {
  msm_shape sh;
  sh.n = n;
  sh.c = c;
  sh.W = (bits + 1 + c - 1) / c;
  sh.B = 1u << (c - 1);
  if (factor < 1) factor = 1;
  if (batch < 1) batch = 1;
  sh.Wf = (sh.W + factor - 1) / factor;
  sh.Bs = sh.B + 1;
  sh.nbuckets = batch * sh.Wf * sh.Bs;
  sh.f = factor;
  sh.batch = batch;
  sh.shared = shared ? 1u : 0u;
  sh.chunk_log = 31;
  sh.nchunks = 1;
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

// ---------------------------------------------------------------- 1 digits, 2 grouping by bucket
// Signed-digit recoding with the same digit set as the reference
// (msm_kernels.cu:96-130): d in [-(2^(c-1)-1), 2^(c-1)], d > 2^(c-1) => d -= 2^c, carry.
// Digit w of canonical scalar s: bucket slot `key` = wf*Bs + |d|-1 (zero digits: the set's trash slot wf*Bs + B) and
// entry `val` = (base index << 1) | sign.  `carry` runs from window 0 upwards.
// i = index of the scalar within its MSM, b = which MSM of the batch.
B381_DI void msm_digit_at(const fr_t& s, const msm_shape& sh, uint32_t i, uint32_t b, uint32_t w, uint32_t& carry,
                          uint32_t& key, uint32_t& val) {
  const uint32_t bit = w * sh.c, limb = bit >> 6, off = bit & 63;
  uint32_t d = 0;
  if (limb < 4) {
    uint64_t lo = s.l[limb] >> off;
    if (off + sh.c > 64 && limb + 1 < 4) lo |= s.l[limb + 1] << (64 - off);
    d = (uint32_t)lo & ((1u << sh.c) - 1u);
  }
  d += carry;
  carry = 0;
  uint32_t sign = 0;
  if (d > sh.B) {          // B = 2^(c-1)
    d = (1u << sh.c) - d;
    sign = 1;
    carry = 1;
  }
  const uint32_t blk = w / sh.Wf, wf = w - blk * sh.Wf;
  key = (b * sh.Wf + wf) * sh.Bs + (d ? d - 1 : sh.B);
  val = (((sh.shared ? i : b * sh.n + i) * sh.f + blk) << 1) | sign;
  // W*c >= 256 > bit length of any canonical scalar, so the last carry is always 0.
}

// counter += v, returning the old value: an L2 atomic on the device (RED when the result is unused), a plain
// increment in the single-threaded host simulation
B381_DI uint32_t msm_fetch_add(uint32_t* p, uint32_t v) {
#ifdef B381_HOST_TEST
  uint32_t o = *p;
  *p = o + v;
  return o;
#else
  return atomicAdd(p, v);
#endif
}

// The (bucket, point) pairs are grouped by a counting sort, hand-written because the keys are dense small integers
// (nbuckets = Wf * (2^(c-1) + 1) slots) and only the GROUPING matters, not the order inside a bucket -- the bucket
// sum is the same group element whatever the order, and the MSM result is its unique affine form:
//   pass 1 (this body)   hist[run]++ for every (scalar, window)            -- 2^28 REDs at 2^24 points, 1.2 ms on B200
//                        (run: see CHUNK-MAJOR GROUPING above; = key when there is one chunk)
//   scan                 offsets = exclusive scan of hist (msm_sort.cu)    -- these ARE the bucket boundaries
//   pass 2 (next body)   vals[cursor[key]++] = val                         -- digits recomputed, no key array exists
// The reference writes 2 x 32-bit keys/values per pair and radix-sorts all 32 key bits with CUB
// (msm_kernels.cu:69-143, :768-778), then histograms the sorted keys with atomics (:224-256).
// t = position in the batch-major scalar array [batch][n]
B381_DI void msm_hist_body(uint32_t t, const fr_t* scalars, bool scalars_mont, const msm_shape sh, uint32_t* hist) {
  if (t >= sh.n * sh.batch) return;
  const uint32_t b = t / sh.n, i = t - b * sh.n;
  fr_t s = scalars[t];
  if (scalars_mont) s = from_mont(s);
  uint32_t carry = 0, key, val;
  for (uint32_t w = 0; w < sh.W; w++) {
    msm_digit_at(s, sh, i, b, w, carry, key, val);
    msm_fetch_add(hist + msm_run_of(sh, b, i, key), 1u);
  }
}

// pass 2: eight windows at a time so that eight returning atomics, then eight stores, are in flight per thread
B381_DI void msm_scatter_body(uint32_t t, const fr_t* scalars, bool scalars_mont, const msm_shape sh, uint32_t* cursor,
                              uint32_t* vals) {
  if (t >= sh.n * sh.batch) return;
  const uint32_t b = t / sh.n, i = t - b * sh.n;
  fr_t s = scalars[t];
  if (scalars_mont) s = from_mont(s);
  uint32_t carry = 0;
  for (uint32_t w0 = 0; w0 < sh.W; w0 += 8) {
    uint32_t key[8], val[8], pos[8];
#ifndef B381_HOST_TEST
#pragma unroll
#endif
    for (uint32_t j = 0; j < 8; j++)
      if (w0 + j < sh.W) msm_digit_at(s, sh, i, b, w0 + j, carry, key[j], val[j]);
#ifndef B381_HOST_TEST
#pragma unroll
#endif
    for (uint32_t j = 0; j < 8; j++)
      if (w0 + j < sh.W) pos[j] = msm_fetch_add(cursor + msm_run_of(sh, b, i, key[j]), 1u);
#ifndef B381_HOST_TEST
#pragma unroll
#endif
    for (uint32_t j = 0; j < 8; j++)
      if (w0 + j < sh.W) vals[pos[j]] = val[j];
  }
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

// Visiting order of the tasks: longest first, so the 32 tasks of a warp have near-equal lengths and the grid's tail
// is made of the shortest ones.  Counting sort over the K + 1 keys K - len (msm_sort.cu); capped buckets have tasks
// longer than K: key 0, visited first.
B381_DI uint32_t msm_task_key(uint32_t t, const uint2* tasks, uint32_t K) {
  const uint2 tk = tasks[t];
  const uint32_t len = tk.y - tk.x;
  return len < K ? K - len : 0u;
}

// ---------------------------------------------------------------- 5 accumulate (hot)
// `pts` is either the caller's bases, visited through the sorted entries, or the output of the last affine level
// (msm_batch.cuh level_pts: two coordinate arrays, entry j at position j, sign already applied).
template <class F>
B381_DI void msm_accumulate_body(uint32_t t, uint32_t ntasks, const uint2* tasks, const uint32_t* sorted_vals,
                                 const level_pts<F>& pts, xyzz_t<F>* partial, const uint32_t* order = nullptr) {
  if (t >= ntasks) return;
  if (order) t = order[t];              // tasks visited longest-first so a warp's 32 tasks have similar lengths
  uint2 tk = tasks[t];
  xyzz_t<F> acc = xyzz_identity<F>();
  for (uint32_t j = tk.x; j < tk.y; j++) {
    const uint32_t v = pts.aos ? sorted_vals[j] : 0u;
    affine_t<F> p = pts.aos ? pts.aos[v >> 1] : affine_t<F>{load_vec16(pts.x + j), load_vec16(pts.y + j)};
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
