#pragma once
// Pippenger MSM for BLS12-381 G1 and G2 on sm_100a: kernels + host driver, templated on the coordinate
// field.  Instantiated by msm_g1.cu (Fq) and msm_g2.cu (Fq2) -- two translation units so the two
// groups compile in parallel; the C ABI lives there.
// Kernel bodies live in msm_core.cuh (host-testable); pipeline description there.
//
// Replaces: msm::msm_cuda<S,A,P> (bls12-381/src/curve/msm_kernels.cu:603-903) and the ICICLE
// wrappers msm_cuda_impl / msm_g2_cuda_impl / msm_precompute_bases_cuda_impl
// (bls12-381/src/backend/icicle_curve_api.cu:243-407, :415-440, :454-650).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "msm_core.cuh"
#include "msm_sort.cuh"

namespace b381 {

// ------------------------------------------------------------------ kernels
static __global__ void k_msm_task_count(const uint32_t* offsets, uint32_t nbuckets, uint32_t Bs, uint32_t K,
                                        uint32_t* counts) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == nbuckets) counts[b] = 0;  // sentinel so the exclusive scan yields the task total
  msm_task_count_body(b, offsets, nbuckets, Bs, K, counts);
}

static __global__ void k_msm_build_tasks(const uint32_t* offsets, const uint32_t* task_start, uint32_t nbuckets,
                                         uint32_t Bs, uint32_t K, uint2* tasks) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  msm_build_tasks_body(b, offsets, task_start, nbuckets, Bs, K, tasks);
}

template <class F, int MINB = 1>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate(const uint32_t* ntasks_dev, const uint2* tasks,
                                                              const uint32_t* sorted_vals, const level_pts<F> pts,
                                                              xyzz_t<F>* partial, const uint32_t* order) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  msm_accumulate_body<F>(t, *ntasks_dev, tasks, sorted_vals, pts, partial, order);
}

// ---- affine pre-reduction (msm_batch.cuh): three kernels per level, no barriers ----------------
static __global__ void k_msm_half_counts(const uint32_t* offsets, uint32_t nbuckets, uint32_t Bs, uint32_t* counts) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  msm_half_counts_body(b, offsets, nbuckets, Bs, counts);
}

// runs [r_begin, r_end] (r_end = sentinel); the whole level: r_begin = 0, r_end >= the number of runs
static __global__ void k_msm_half_counts_runs(const uint32_t* run_off, uint32_t nbuckets, uint32_t set_slots, uint32_t nchunks,
                                              uint32_t Bs, uint32_t* counts, uint32_t* counts_bucket_major,
                                              uint32_t r_begin = 0u, uint32_t r_end = 0xFFFFFFFFu) {
  msm_half_counts_runs_body(r_begin + blockIdx.x * blockDim.x + threadIdx.x, run_off, nbuckets, set_slots, nchunks, Bs,
                            counts, counts_bucket_major, r_end);
}
// bucket slot k's level-1 entries start where its chunk-0 run's sums were written
static __global__ void k_msm_level1_offsets(const uint32_t* dst_base, uint32_t nbuckets, uint32_t nchunks, uint32_t* off1) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= nbuckets) off1[k] = dst_base[(size_t)k * nchunks];
}

// The three kernels of a level live in msm_pair.cu (their own translation unit: seconds to compile).
// Launches k_msm_pair_fwd, k_msm_invert_totals, k_msm_pair_bwd for one level; `grid` CTAs of PR_TPB threads,
// nt = grid * PR_TPB = stride of the slot-major scratch arrays.
// msm_pair_levels: how many levels pay for `total` sorted entries at a mean bucket load of `avg`.
int msm_pair_levels(double avg, size_t total);
template <class F>
void launch_pair_level(bool level0, const uint32_t* in_off, const uint32_t* out_off, uint32_t nbuckets,
                       const uint32_t* svals, const level_pts<F> pts, size_t npts, unsigned grid, uint32_t* srcg, F* preg,
                       F* tot, F* outx, F* outy, cudaStream_t st, const uint32_t* dst_base = nullptr,
                       uint32_t dst_set_slots = 0, uint32_t dst_nchunks = 0, uint32_t* dst_slots = nullptr);

// Streamed chunk-major level 0 (plugin call with host scalars): forward pass of one piece of runs while the next piece
// is still crossing PCIe, then destinations + inversion + backward pass once every piece is in (msm_pair.cu)
template <class F>
void launch_pair_fwd_piece(const uint32_t* in_off, const uint32_t* out_off, const uint32_t* svals, const level_pts<F> pts,
                           unsigned grid, unsigned piece_grid, uint32_t t0, uint32_t nb_search, bool final,
                           const uint32_t* lo_dev, const uint32_t* hi_dev, uint32_t* srcg, F* preg, F* tot, cudaStream_t st);
template <class F>
void launch_pair_finish_streamed(const uint32_t* in_off, const uint32_t* out_off, uint32_t nbuckets, const uint32_t* svals,
                                 const level_pts<F> pts, unsigned grid, uint32_t* srcg, F* preg, F* tot, F* outx, F* outy,
                                 cudaStream_t st, const uint32_t* dst_base, uint32_t dst_set_slots, uint32_t dst_nchunks,
                                 uint32_t* dst_slots);

// window sums from the buckets, four lanes per segment: msm_tail.cu
template <class F>
void launch_msm_bucket_reduce(uint32_t W, uint32_t B, uint32_t L, const xyzz_t<F>* buckets, xyzz_t<F>* seg, cudaStream_t st);
// bucket = sum of its task partials (thread per bucket, warp per listed heavy bucket): msm_tail.cu
template <class F>
void launch_msm_finalize(uint32_t nbuckets, const uint32_t* task_start, const uint32_t* counts, const xyzz_t<F>* partial,
                         xyzz_t<F>* buckets, uint32_t* heavy_list, uint32_t* heavy_count, cudaStream_t st);

// out[b] = Horner over the W window sums of MSM b (wsum + b*W*stride), one CTA per MSM of the batch
template <class F>
void launch_msm_combine(const xyzz_t<F>* wsum, uint32_t stride, uint32_t W, uint32_t c, xyzz_t<F>* out, uint32_t batch,
                        cudaStream_t st);

template <class F>
__global__ void k_msm_set_identity(xyzz_t<F>* out, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = xyzz_identity<F>();
}

// out[r] = ICICLE encoding of the sum of the `count` XYZZ partials parts[r*count ..], one thread per result
template <class F>
__global__ void k_msm_encode(const xyzz_t<F>* parts, int count, bool mont, jacobian_t<F>* out, uint32_t results) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= results) return;
  xyzz_t<F> acc = xyzz_identity<F>();
  for (int i = 0; i < count; i++) xyzz_add(acc, parts[(size_t)r * count + i]);
  out[r] = msm_result_encode<F>(acc, mont);
}

template <class F>
__global__ void k_points_to_mont(const affine_t<F>* in, affine_t<F>* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  affine_t<F> p = in[i];
  out[i] = affine_t<F>{to_mont(p.x), to_mont(p.y)};
}

// out[i*factor + k] = 2^(k*shift_bits) * in[i], one thread per input point.  `in` is Montgomery affine (the entry
// point converted it if needed); `out` is written in the form the caller declared for the input, like upstream
// ICICLE's precompute_bases does.
template <class F>
__global__ void __launch_bounds__(64) k_precompute_bases(const affine_t<F>* in, affine_t<F>* out, uint32_t n,
                                                         uint32_t factor, uint32_t shift_bits, bool out_mont) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  affine_t<F> p = in[i];
  xyzz_t<F> acc = to_xyzz(p);
  for (uint32_t k = 0; k < factor; k++) {
    if (k) {
      for (uint32_t b = 0; b < shift_bits; b++) acc = xyzz_dbl(acc);
      p = xyzz_to_affine(acc);
    }
    out[(size_t)i * factor + k] = out_mont ? p : affine_t<F>{from_mont(p.x), from_mont(p.y)};
  }
}

// Which form are these points in?  The reference's Rust layer declares precomputed bases Montgomery when it
// builds them and NOT Montgomery when it uses them (core/msm.rs:450 vs :641-643), so for precompute_factor > 1
// the flag cannot be trusted: the first finite point is tested against y^2 = x^3 + b read as Montgomery words.
// A point satisfies the equation in both readings with probability ~2^-381.  *flag = 1 Montgomery, 0 not, 2 = no
// finite point among the first `count` (either reading gives the identity).
template <class F>
__global__ void k_points_form_probe(const affine_t<F>* pts, uint32_t count, int* flag) {
  if (blockIdx.x || threadIdx.x) return;
  *flag = 2;
  for (uint32_t i = 0; i < count; i++) {
    affine_t<F> p = pts[i];
    if (is_inf(p)) continue;
    *flag = on_curve(p) ? 1 : 0;
    return;
  }
}

// ------------------------------------------------------------------ host driver
inline thread_local float g_last_timings[12];
inline thread_local int g_last_timings_n = 0;
inline thread_local int g_last_info[4] = {0, 0, 0, 0};   // window bits, windows, affine levels, own kernel launches

struct PhaseTimer {
  bool on;
  cudaStream_t s;
  std::vector<cudaEvent_t> ev;
  PhaseTimer(cudaStream_t st) : s(st) {
    const char* e = getenv("B381_MSM_TIMING");
    on = e && e[0] == '1';
  }
  // phase boundary: closes the NVTX range of the phase that ends here and opens `next` (nullptr: none)
  bool range_open = false;
  void mark(const char* next = nullptr) {
    if (range_open) nvtxDomainRangePop(TraceRange::domain());
    range_open = next != nullptr;
    if (next) {
      nvtxEventAttributes_t a = {};
      a.version = NVTX_VERSION; a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
      a.messageType = NVTX_MESSAGE_TYPE_ASCII; a.message.ascii = next;
      nvtxDomainRangePushEx(TraceRange::domain(), &a);
    }
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
  }
  ~PhaseTimer() { if (range_open) nvtxDomainRangePop(TraceRange::domain()); }
  void finish() {
    if (!on || ev.empty()) return;
    cudaEventSynchronize(ev.back());
    g_last_timings_n = 0;
    for (size_t i = 1; i < ev.size() && g_last_timings_n < 12; i++)
      cudaEventElapsedTime(&g_last_timings[g_last_timings_n++], ev[i - 1], ev[i]);
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

static uint32_t ceil_log2_u64(uint64_t v) {
  uint32_t b = 0;
  while ((1ull << b) < v) b++;
  return b;
}

// Window size (replaces get_optimal_c, include/msm.cuh:115-140).  Full-width G1 scalars without precomputed-bases
// folding use the optimum MEASURED on B200 with scalars uniform in [0, r) (gpurun sweep of c = 12..16 at 2^12..2^21,
// profiles/r01c_msm_window_sweep.txt): c = 13 up to 2^17 points (2^16: 2.80 ms vs 3.24 at c = 16), c = 16 above
// (2^20: 9.07 ms vs 9.87 at c = 13).  The tail (bucket reduction + window combine) is latency- not throughput-bound,
// so few large windows win much earlier than a product count predicts.  c = 15 is never the optimum for uniform
// scalars: 17 windows cover exactly 255 bits, so an 18th window exists only for the signed-digit carry and its single
// bucket receives 45 % of the points (r >> 240 = 0x73ed > 2^14).  At c = 16 the top digit is at most 0x73ed < 2^15 and
// never carries.  Everything else (short scalars, folded windows, G2) minimises
//   W * 10 n  +  Wf * 72 * 2^(c-1)      [Fq-mul equivalents: bucket insertions + bucket reduction]
// with W = ceil((bits+1)/c), capped at c = 16: beyond that the per-bucket load (n / 2^(c-1)) gets so
// small that warp divergence in the accumulate kernel eats the saving (measured on B200: 2^24 points,
// c = 16 -> 130 ms, c = 20 -> 191 ms; profiles/r01_msm_window_sweep.txt).
constexpr uint32_t kPrecomputeWindow = 16;   // cfg.c == 0 with precompute_factor > 1
static uint32_t pick_window(uint32_t n, uint32_t bits, uint32_t factor, bool g1 = true) {
  const char* e = getenv("B381_MSM_C");
  if (e && atoi(e) > 0) return (uint32_t)atoi(e);
  // Precomputed bases: the stored multiples are 2^(k*c*Wf) P, so c must be the SAME when the table is built and every
  // time it is used, whatever prefix length n the MSM runs over -- it may not depend on n.
  if (factor > 1) return kPrecomputeWindow;
  if (g1 && factor <= 1 && bits >= 250 && n >= (1u << 11)) return n <= (1u << 17) ? 13u : 16u;
  double best = 1e300;
  uint32_t bc = 4;
  for (uint32_t c = 4; c <= 16; c++) {
    uint32_t W = (bits + 1 + c - 1) / c;
    uint32_t Wf = (W + factor - 1) / factor;
    // the top window holds only t = bits+1 - (W-1)c bits: with t << c its 2^t buckets each take n/2^t
    // points (n = 2^16, c = 11: five buckets of 16384) -- keep it within 8x of the other windows' load
    uint32_t t = bits + 1 - (W - 1) * c;
    if (t + 4 < c) continue;
    double cost = (double)W * 10.0 * n + (double)Wf * 72.0 * (double)(1u << (c - 1));
    if (cost < best) { best = cost; bc = c; }
  }
  return bc;
}

template <class F>
// `batch` MSMs of n points each in ONE pipeline run (msm_core.cuh make_msm_shape): scalars [batch][n], bases shared or
// [batch][n * factor], d_out[batch].
static cudaError_t msm_single(Scratch& sc, const fr_t* d_scalars, bool scalars_mont, const affine_t<F>* d_bases,
                              uint32_t n, uint32_t c_req, uint32_t bits, uint32_t factor, xyzz_t<F>* d_out,
                              const fr_t* host_scalars = nullptr, uint32_t batch = 1, bool shared = true) {
  cudaStream_t st = sc.stream();
  if (n == 0) {
    k_msm_set_identity<F><<<grid_for(batch, 32), 32, 0, st>>>(d_out, batch);
    return cudaGetLastError();
  }
  PhaseTimer tm(st);
  uint32_t c = c_req ? c_req : pick_window(n, bits, factor, sizeof(F) == sizeof(fq_t));
  if (c < 2) c = 2;
  if (c > 24) c = 24;
  msm_shape sh = make_msm_shape(n, c, bits, factor, batch, shared);
  const size_t npts = (size_t)n * factor * (shared ? 1u : batch);
  if (npts >= (1ull << 31) || (uint64_t)sh.Wf * sh.Bs * batch >= (1ull << 31)) return cudaErrorInvalidValue;
  size_t total = (size_t)n * sh.W * batch;
  if (total >= (1ull << 31)) return cudaErrorInvalidValue;

  // how many affine levels will run (a level pays once there are enough pairs to fill the GPU and buckets are still long)
  double avg = (double)n * sh.W / ((double)sh.Wf * sh.B);   // mean bucket load seen by the accumulate kernel
  int levels = msm_pair_levels(avg, total);
  {
    const char* e = getenv("B381_MSM_LEVELS");
    if (e && e[0]) levels = atoi(e);
    if (levels > 16) levels = 16;
  }
  // chunk-major grouping (msm_core.cuh) when level 0 runs and the bases of ONE MSM are several times larger than the L2
  // can hold: chunks of 100 MB of bases (2^20 G1 / 2^19 G2 points), from 8 chunks up, at most 64.  Measured on B200
  // (profiles/r02d_msm_chunk_sweep.txt): 2^24 points 77.5 -> 74.3 ms (level-0 forward pass 10.7 -> 7.0 ms), 2^23
  // 40.7 -> 39.6; chunks of 2^19 / 2^18 are slower (shorter runs: more carried-over singles), 2^22 points gain nothing.
  {
    const bool g1 = sizeof(F) == sizeof(fq_t);
    uint32_t chunk_log = g1 ? 20u : 19u;
    bool worth = levels >= 1 && (uint64_t)n >= (8ull << chunk_log);
    // Host scalars (the plugin call): chunk-major also pays from FOUR chunks up for G1, because chunks are what lets the
    // sort and level 0's forward pass run under the PCIe copy (msm_sort.cu pieces, `streamed` below).  Measured on B200
    // (profiles/r02h_host_chunks.txt): 2^22 host scalars 23.96 -> 22.75 ms; 2^21 in four 2^19 chunks 13.78 -> 13.84 and
    // G2 2^20 / 2^21 24.4 -> 25.1 / 40.5 -> 40.8 (no gain: not enabled); resident scalars gain nothing below 2^23.
    if (!worth && g1 && host_scalars && batch == 1 && levels >= 1 && (uint64_t)n >= (4ull << chunk_log)) worth = true;
    if (const char* e = getenv("B381_MSM_CHUNK_LOG")) {          // A/B runs and tests: any size, 31 = off
      chunk_log = (uint32_t)atoi(e);
      worth = levels >= 1 && chunk_log < 31 && (uint64_t)n > (1ull << chunk_log);
    }
    while (chunk_log < 31 && (((uint64_t)n + (1ull << chunk_log) - 1) >> chunk_log) > 64) chunk_log++;
    if (worth) msm_shape_set_chunks(sh, chunk_log);
    if ((uint64_t)sh.nchunks * sh.nbuckets >= (1ull << 31)) msm_shape_set_chunks(sh, 31);
  }
  const size_t nruns = msm_runs(sh);

  // plan of the affine levels.  The slot-major scratch (stride nt = grid * PR_TPB) and the ping-pong point buffers are
  // allocated once, for the LARGEST level.  That is level 0 while total > nbuckets + 2, but the bound
  // (in + nbuckets) / 2 + 1 GROWS from level to level on a sparse input (forced B381_MSM_LEVELS with few points per
  // bucket), so take the maximum over the plan.  (Level 0 of a chunk-major run rounds up once per RUN, not once per
  // bucket slot.)
  constexpr int PB = pair_batch<F>::B;
  size_t nt_max = 0, nt_buf[2] = {0, 0};        // scratch stride; point buffers of the even / odd levels
  {
    size_t in = total;
    for (int l = 0; l < levels; l++) {
      const size_t out = (in + (l == 0 ? nruns : (size_t)sh.nbuckets)) / 2 + 1;
      const size_t nt_l = (size_t)grid_for(out, (size_t)PR_TPB * PB) * PR_TPB;
      if (nt_l > nt_max) nt_max = nt_l;
      if (nt_l > nt_buf[l & 1]) nt_buf[l & 1] = nt_l;
      in = out;
    }
  }
  F* bufx[2] = {nullptr, nullptr};              // ping-pong level outputs, struct of arrays (msm_batch.cuh level_pts)
  F* bufy[2] = {nullptr, nullptr};
  uint32_t *srcg = nullptr, *dstg = nullptr;
  F *preg = nullptr, *tot = nullptr;
  auto alloc_level_scratch = [&](int l) -> cudaError_t {
    if (!bufx[l & 1]) {
      B381_CUDA_TRY(sc.alloc(&bufx[l & 1], nt_buf[l & 1] * PB));
      B381_CUDA_TRY(sc.alloc(&bufy[l & 1], nt_buf[l & 1] * PB));
    }
    if (!srcg) {
      B381_CUDA_TRY(sc.alloc(&srcg, nt_max * PB));
      B381_CUDA_TRY(sc.alloc(&preg, nt_max * PB));
      B381_CUDA_TRY(sc.alloc(&tot, nt_max));
      if (sh.nchunks > 1) B381_CUDA_TRY(sc.alloc(&dstg, nt_max * PB));
    }
    return cudaSuccess;
  };

  // -- 1 histogram of the runs, 2 scan = run boundaries, 3 scatter (msm_sort.cu; no library sort)
  int launches = 0;
  uint32_t *vals, *hist, *run_off, *offsets, *counts, *task_start;
  B381_CUDA_TRY(sc.alloc(&vals, total));
  B381_CUDA_TRY(sc.alloc(&hist, nruns + 1));
  B381_CUDA_TRY(sc.alloc(&run_off, nruns + 1));
  B381_CUDA_TRY(sc.alloc(&counts, (size_t)sh.nbuckets + 1));
  B381_CUDA_TRY(sc.alloc(&task_start, (size_t)sh.nbuckets + 1));
  if (sh.nchunks > 1) B381_CUDA_TRY(sc.alloc(&offsets, (size_t)sh.nbuckets + 1));
  else offsets = run_off;                       // one chunk: runs ARE bucket slots

  // Streamed level 0 (host scalars, chunk-major, one MSM, several pieces): a piece's runs are final as soon as the
  // piece is scattered, so its share of the level-0 FORWARD pass (slot walk, x-gathers, prefix products) runs while the
  // next piece is still crossing PCIe.  Slot numbers continue from piece to piece through a device word per piece
  // (carry[p] = slots before piece p); a thread runs in the piece that completes its PB slots.  What needs every piece
  // -- the bucket-major destinations, the batched inversion, the backward pass -- follows the last one.
  static const bool stream_l0_off = [] { const char* e = getenv("B381_MSM_STREAM_L0"); return e && e[0] == '0'; }();
  const bool streamed = levels >= 1 && sh.nchunks > 1 && !stream_l0_off && msm_sort_is_streamed(sh, host_scalars);
  uint32_t *half0 = nullptr, *out_off0 = nullptr, *half_bm0 = nullptr, *dst_base0 = nullptr, *carry = nullptr;
  const unsigned g0 = grid_for((total + nruns) / 2 + 1, (size_t)PR_TPB * PB);
  const level_pts<F> base_pts = level_from_bases(d_bases);
  uint32_t piece_no = 0;
  msm_piece_fn after_piece = [&](size_t r0, size_t r1, bool last) -> cudaError_t {
    const size_t cnt = r1 - r0 + 1;             // + the sentinel: the scan leaves the slot total up to r1 at out_off0[r1]
    k_msm_half_counts_runs<<<grid_for(cnt, 256), 256, 0, st>>>(run_off, sh.nbuckets, msm_set_slots(sh), sh.nchunks, sh.Bs,
                                                               half0, half_bm0, (uint32_t)r0, (uint32_t)r1);
    B381_CUDA_TRY(exclusive_scan_u32(sc, half0 + r0, out_off0 + r0, cnt, nullptr, &launches, 0u, carry + piece_no,
                                     carry + piece_no + 1));
    // the host knows only an upper bound of the slots so far: one per two entries + one per run
    uint64_t pts_end = (uint64_t)(r1 / sh.nbuckets) << sh.chunk_log;
    if (pts_end > n) pts_end = n;
    unsigned pg = grid_for((size_t)((pts_end * sh.W + r1) / 2 + 1), (size_t)PR_TPB * PB);
    if (pg > g0 || last) pg = g0;
    launch_pair_fwd_piece<F>(run_off, out_off0, vals, base_pts, g0, pg, 0u, (uint32_t)r1, last, carry + piece_no,
                             carry + piece_no + 1, srcg, preg, tot, st);
    launches += 2;
    piece_no++;
    return cudaGetLastError();
  };
  if (streamed) {
    const size_t npieces = ((size_t)n + msm_sort_piece(sh) - 1) / msm_sort_piece(sh);
    B381_CUDA_TRY(sc.alloc(&half0, nruns + 1));
    B381_CUDA_TRY(sc.alloc(&out_off0, nruns + 1));
    B381_CUDA_TRY(sc.alloc(&half_bm0, nruns + 1));
    B381_CUDA_TRY(sc.alloc(&dst_base0, nruns + 1));
    B381_CUDA_TRY(sc.alloc(&carry, npieces + 1));
    B381_CUDA_TRY(alloc_level_scratch(0));
    B381_CUDA_TRY(cudaMemsetAsync(carry, 0, sizeof(uint32_t), st));
  }
  tm.mark("msm:sort");
  B381_CUDA_TRY(msm_sort_pairs(sc, d_scalars, scalars_mont, sh, host_scalars, hist, run_off, vals, &launches,
                               streamed ? &after_piece : nullptr));
  const uint32_t* svals = vals;
  tm.mark(nullptr);
  tm.mark(nullptr);     // (two phase slots kept: histogram / scatter / offsets were separately timed passes before)
  tm.mark("msm:affine_levels");

  // -- 3b affine pre-reduction levels (msm_batch.cuh): each halves every bucket
  level_pts<F> acc_pts = base_pts;
  const uint32_t* acc_vals = svals;
  int n_levels = 0;
  {
    size_t max_in = total;
    const uint32_t* in_off = run_off;
    for (int l = 0; l < levels; l++) {
      const bool chunked = l == 0 && sh.nchunks > 1;
      const size_t nb_l = chunked ? nruns : (size_t)sh.nbuckets;        // "buckets" of this level
      const size_t max_out = (max_in + nb_l) / 2 + 1;
      const unsigned g = grid_for(max_out, (size_t)PR_TPB * PB);
      B381_CUDA_TRY(alloc_level_scratch(l));
      uint32_t* dst_base = nullptr;
      const uint32_t* out_off_l;
      if (chunked && streamed) {
        // forward pass done piece by piece under the copy (after_piece above)
        dst_base = dst_base0;
        out_off_l = out_off0;
        B381_CUDA_TRY(exclusive_scan_u32(sc, half_bm0, dst_base0, nb_l + 1, nullptr, &launches));
        launch_pair_finish_streamed<F>(in_off, out_off0, (uint32_t)nb_l, svals, acc_pts, g, srcg, preg, tot, bufx[0], bufy[0],
                                       st, dst_base0, msm_set_slots(sh), sh.nchunks, dstg);
        launches += 3;    // destinations, invert, backward
      } else {
        uint32_t *half, *out_off;
        B381_CUDA_TRY(sc.alloc(&half, nb_l + 1));
        B381_CUDA_TRY(sc.alloc(&out_off, nb_l + 1));
        if (chunked) {
          // sums are produced run by run (chunk-major) and WRITTEN bucket-major: second count array in (slot, chunk) order
          uint32_t* half_bm;
          B381_CUDA_TRY(sc.alloc(&half_bm, nb_l + 1));
          B381_CUDA_TRY(sc.alloc(&dst_base, nb_l + 1));
          k_msm_half_counts_runs<<<grid_for(nb_l + 1, 256), 256, 0, st>>>(in_off, sh.nbuckets, msm_set_slots(sh), sh.nchunks,
                                                                          sh.Bs, half, half_bm);
          B381_CUDA_TRY(exclusive_scan_u32(sc, half_bm, dst_base, nb_l + 1, nullptr, &launches));
        } else {
          k_msm_half_counts<<<grid_for(nb_l + 1, 256), 256, 0, st>>>(in_off, sh.nbuckets, sh.Bs, half);
        }
        B381_CUDA_TRY(exclusive_scan_u32(sc, half, out_off, nb_l + 1, nullptr, &launches));
        launches += 4;      // half counts + forward, invert, backward
        out_off_l = out_off;
        launch_pair_level<F>(l == 0, in_off, out_off, (uint32_t)nb_l, l == 0 ? svals : nullptr, acc_pts, npts, g, srcg, preg,
                             tot, bufx[l & 1], bufy[l & 1], st, dst_base, msm_set_slots(sh), sh.nchunks, dstg);
      }
      acc_pts = level_from_xy<F>(bufx[l & 1], bufy[l & 1]);
      acc_vals = nullptr;
      if (chunked) {
        // from here on nothing knows about chunks: bucket slot k's entries are the sums of its runs, chunk after chunk
        uint32_t* off1;
        B381_CUDA_TRY(sc.alloc(&off1, (size_t)sh.nbuckets + 1));
        k_msm_level1_offsets<<<grid_for((size_t)sh.nbuckets + 1, 256), 256, 0, st>>>(dst_base, sh.nbuckets, sh.nchunks, off1);
        launches++;
        in_off = off1;
      } else {
        in_off = out_off_l;
      }
      max_in = max_out;
      avg *= 0.5;
    }
    n_levels = levels;
    if (levels) {
      // the task builder and the finalize step read bucket boundaries from `offsets`
      B381_CUDA_TRY(cudaMemcpyAsync(offsets, in_off, sizeof(uint32_t) * ((size_t)sh.nbuckets + 1), cudaMemcpyDeviceToDevice, st));
      total = max_in;
    }
  }
  tm.mark("msm:tasks+accumulate");

  // -- 4 tasks
  // task length bound: mean bucket load + 4 sigma (Poisson), so a uniform input is one task per bucket
  if (avg < 1.0) avg = 1.0;
  uint32_t K = (uint32_t)(avg + 4.0 * sqrt(avg) + 8.0);
  {
    const char* e = getenv("B381_MSM_K");
    if (e && atoi(e) > 0) K = (uint32_t)atoi(e);
  }
  k_msm_task_count<<<grid_for((size_t)sh.nbuckets + 1, 256), 256, 0, st>>>(offsets, sh.nbuckets, sh.Bs, K, counts);
  B381_CUDA_TRY(exclusive_scan_u32(sc, counts, task_start, (size_t)sh.nbuckets + 1, nullptr, &launches));
  const size_t max_tasks = (size_t)sh.nbuckets + total / K + 1;
  uint2* tasks;
  B381_CUDA_TRY(sc.alloc(&tasks, max_tasks));
  k_msm_build_tasks<<<grid_for(sh.nbuckets, 256), 256, 0, st>>>(offsets, task_start, sh.nbuckets, sh.Bs, K, tasks);
  // visiting order: longest task first (counting sort by task length, msm_sort.cu), so the 32 tasks of a warp have
  // near-equal lengths and the grid's tail is made of the shortest ones
  uint32_t* order = nullptr;
  {
    const char* e = getenv("B381_MSM_NO_TASK_SORT");
    if (!(e && e[0] == '1')) {
      B381_CUDA_TRY(sc.alloc(&order, max_tasks));
      B381_CUDA_TRY(msm_task_order(sc, max_tasks, task_start + sh.nbuckets, tasks, K, order, &launches));
    }
  }

  // -- 5 accumulate
  xyzz_t<F>*partial, *buckets;
  B381_CUDA_TRY(sc.alloc(&partial, max_tasks));
  B381_CUDA_TRY(sc.alloc(&buckets, (size_t)sh.nbuckets));
  {
    // G1: 166 registers -> 3 CTAs (12 warps) per SM: best of {1,3,4} on B200 (119.4 / 116.0 / 121.8 ms at 2^24).  G2: at that
    // cap the Fq2 state spills 808 bytes per thread; uncapped (255 registers) 2^20: 2.88 -> 2.44 ms (profiles/r02h_g2_occupancy.txt)
    int variant = sizeof(F) > sizeof(fq_t) ? 1 : 3;
    const char* e = getenv("B381_ACC_MINB");
    if (e) variant = atoi(e);
    const unsigned g = grid_for(max_tasks, 128);
    if (variant == 3) k_msm_accumulate<F, 3><<<g, 128, 0, st>>>(task_start + sh.nbuckets, tasks, acc_vals, acc_pts, partial, order);
    else if (variant == 4) k_msm_accumulate<F, 4><<<g, 128, 0, st>>>(task_start + sh.nbuckets, tasks, acc_vals, acc_pts, partial, order);
    else k_msm_accumulate<F, 1><<<g, 128, 0, st>>>(task_start + sh.nbuckets, tasks, acc_vals, acc_pts, partial, order);
  }
  tm.mark("msm:finalize");
  // -- 6 finalize
  {
    uint32_t* heavy;           // [0] = count, [1 ..] = list
    B381_CUDA_TRY(sc.alloc(&heavy, (size_t)sh.nbuckets + 1));
    launch_msm_finalize<F>(sh.nbuckets, task_start, counts, partial, buckets, heavy + 1, heavy, st);
  }
  tm.mark("msm:bucket_reduce");

  // -- 7 segments, 8 tree
  // segment length: measured on B200 (gpurun sweep of B381_MSM_L = 4..64): 32 up to 2^14 buckets per window
  // (2^21 points: 0.73 ms vs 0.97 at 64, 0.94 at 16), 64 at 2^15 (2^24 points: 1.23 ms vs 1.43) where the kernel is
  // throughput- rather than latency-bound and the per-segment scalar multiplication is amortised over more buckets
  uint32_t L = sh.B >= (1u << 15) ? 64 : 32;
  {
    const char* e = getenv("B381_MSM_L");
    if (e && atoi(e) > 0) L = (uint32_t)atoi(e);
  }
  while (L > sh.B) L >>= 1;
  uint32_t segs = sh.B / L;
  xyzz_t<F>* seg;
  B381_CUDA_TRY(sc.alloc(&seg, (size_t)sh.Wf * batch * segs));
  launch_msm_bucket_reduce<F>(sh.Wf * batch, sh.B, L, buckets, seg, st);
  tm.mark("msm:combine");
  // -- 9 combine
  launch_msm_combine<F>(seg, segs, sh.Wf, sh.c, d_out, batch, st);
  tm.mark(nullptr);
  {
    int tree = 0;
    for (uint32_t half = segs / 2; half >= 1; half >>= 1) tree++;
    // counted above: histogram, scans, scatter, 4 per level + its scan, task order; here: task_count, build_tasks,
    // accumulate, finalize (2) | segment, tree levels, combine
    g_last_info[0] = (int)sh.c; g_last_info[1] = (int)sh.W; g_last_info[2] = n_levels;
    g_last_info[3] = launches + 5 + 1 + tree + 1;
  }
  B381_CUDA_TRY(cudaGetLastError());
  tm.finish();
  return cudaSuccess;
}

enum class ResultKind { IcicleStd, JacobianMont, PartialXyzz };

template <class F>
static int msm_entry(const fr_t* scalars, const affine_t<F>* bases, int msm_size, const b381_msm_config* cfg,
                     void* results, ResultKind kind) {
  if (!cfg || !results) return B381_INVALID_POINTER;
  if (msm_size < 0) return B381_INVALID_ARGUMENT;
  if (msm_size > 0 && (!scalars || !bases)) return B381_INVALID_POINTER;
  int batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
  if (cfg->c < 0 || cfg->c > 24) return B381_INVALID_ARGUMENT;
  if (cfg->bitsize < 0 || cfg->bitsize > 256) return B381_INVALID_ARGUMENT;
  uint32_t bits = cfg->bitsize > 0 ? (uint32_t)cfg->bitsize : 255u;
  uint32_t n = (uint32_t)msm_size;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  TraceRange trace(sizeof(F) == sizeof(fq_t) ? "b381_g1_msm" : "b381_g2_msm", msm_size);
  {
    Scratch sc(st);
    const fr_t* d_scalars = nullptr;
    const affine_t<F>* d_bases = nullptr;
    uint32_t factor = cfg->precompute_factor > 1 ? (uint32_t)cfg->precompute_factor : 1u;
    size_t nbases = (cfg->are_points_shared_in_batch ? (size_t)n : (size_t)n * batch) * factor;
    // host scalars are not staged here: msm_single copies them in chunks under its histogram pass (msm_sort.cu)
    if (cfg->are_scalars_on_device) d_scalars = scalars;
    else {
      fr_t* buf;
      if ((e = sc.alloc(&buf, (size_t)n * batch)) != cudaSuccess) return map_cuda_error(e);
      d_scalars = buf;
    }
    if ((e = stage_in(sc, bases, nbases, cfg->are_points_on_device, &d_bases)) != cudaSuccess)
      return map_cuda_error(e);
    bool bases_mont = cfg->are_points_montgomery_form;
    if (!bases_mont && factor > 1 && nbases) {
      // precomputed bases: see k_points_form_probe (one 4-byte read-back; only on this flag combination)
      int* d_flag;
      int h_flag = 0;
      if ((e = sc.alloc(&d_flag, 1)) != cudaSuccess) return map_cuda_error(e);
      k_points_form_probe<F><<<1, 1, 0, st>>>(d_bases, (uint32_t)(nbases < 64 ? nbases : 64), d_flag);
      if ((e = cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return map_cuda_error(e);
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return map_cuda_error(e);
      bases_mont = h_flag != 0;
    }
    if (!bases_mont && nbases) {
      affine_t<F>* conv;
      if ((e = sc.alloc(&conv, nbases)) != cudaSuccess) return map_cuda_error(e);
      k_points_to_mont<F><<<grid_for(nbases, 128), 128, 0, st>>>(d_bases, conv, (uint32_t)nbases);
      d_bases = conv;
    }
    xyzz_t<F>* d_part;
    if ((e = sc.alloc(&d_part, (size_t)batch)) != cudaSuccess) return map_cuda_error(e);
    // The batch runs as few pipeline passes as fit: a group of g MSMs needs g*n*W < 2^31 sorted entries and about
    // 110 (G1) / 220 (G2) bytes of scratch per entry (entries, two affine level buffers, slot scratch), kept under half
    // of the memory that is free right now.
    uint32_t group = (uint32_t)batch;
    if (n > 0 && batch > 1) {
      const uint32_t c_eff = cfg->c ? (uint32_t)cfg->c : pick_window(n, bits, factor, sizeof(F) == sizeof(fq_t));
      const uint64_t per_msm = (uint64_t)n * make_msm_shape(n, c_eff < 2 ? 2 : c_eff, bits, factor).W;
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)8 << 30;
      const uint64_t by_mem = (uint64_t)(free_b / 2) / (per_msm * (sizeof(F) == sizeof(fq_t) ? 110u : 220u) + 1);
      const uint64_t by_idx = ((1ull << 31) - 1) / (per_msm + 1);
      uint64_t g = by_mem < by_idx ? by_mem : by_idx;
      if (const char* ev = getenv("B381_MSM_BATCH_GROUP")) g = (uint64_t)atoi(ev);
      group = (uint32_t)(g < 1 ? 1 : g > (uint64_t)batch ? (uint64_t)batch : g);
    }
    for (uint32_t b0 = 0; b0 < (uint32_t)batch; b0 += group) {
      const uint32_t g = (uint32_t)batch - b0 < group ? (uint32_t)batch - b0 : group;
      const bool shared = cfg->are_points_shared_in_batch;
      const affine_t<F>* bb = shared ? d_bases : d_bases + (size_t)b0 * n * factor;
      e = msm_single<F>(sc, d_scalars + (size_t)b0 * n, cfg->are_scalars_montgomery_form, bb, n, (uint32_t)cfg->c,
                        bits, factor, d_part + b0, cfg->are_scalars_on_device ? nullptr : scalars + (size_t)b0 * n, g, shared);
      if (e != cudaSuccess) return map_cuda_error(e);
    }
    if (kind == ResultKind::PartialXyzz) {
      // caller's buffer is device memory by contract
      e = cudaMemcpyAsync(results, d_part, sizeof(xyzz_t<F>) * batch, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return map_cuda_error(e);
    } else {
      jacobian_t<F>* d_res;
      bool direct = cfg->are_results_on_device;
      if (direct) d_res = reinterpret_cast<jacobian_t<F>*>(results);
      else if ((e = sc.alloc(&d_res, (size_t)batch)) != cudaSuccess) return map_cuda_error(e);
      k_msm_encode<F><<<grid_for((size_t)batch, 32), 32, 0, st>>>(d_part, 1, kind == ResultKind::JacobianMont, d_res, (uint32_t)batch);
      if (!direct) {
        e = cudaMemcpyAsync(results, d_res, sizeof(jacobian_t<F>) * batch, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return map_cuda_error(e);
      }
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
  }  // scratch released in stream order
  if (!cfg->is_async) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return map_cuda_error(e);
  }
  return B381_SUCCESS;
}

template <class F>
static int precompute_entry(const affine_t<F>* in, int bases_size, const b381_msm_config* cfg, affine_t<F>* out) {
  if (!cfg || !in || !out) return B381_INVALID_POINTER;
  if (bases_size < 0) return B381_INVALID_ARGUMENT;
  uint32_t n = (uint32_t)bases_size;
  uint32_t factor = cfg->precompute_factor > 1 ? (uint32_t)cfg->precompute_factor : 1u;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  {
    Scratch sc(st);
    const affine_t<F>* d_in = nullptr;
    if ((e = stage_in(sc, in, n, cfg->are_points_on_device, &d_in)) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->are_points_montgomery_form && n) {
      affine_t<F>* conv;
      if ((e = sc.alloc(&conv, (size_t)n)) != cudaSuccess) return map_cuda_error(e);
      k_points_to_mont<F><<<grid_for(n, 128), 128, 0, st>>>(d_in, conv, n);
      d_in = conv;
    }
    // results live where cfg says (ICICLE: are_results_on_device)
    affine_t<F>* d_out = out;
    if (!cfg->are_results_on_device) {
      if ((e = sc.alloc(&d_out, (size_t)n * factor)) != cudaSuccess) return map_cuda_error(e);
    }
    uint32_t bits = cfg->bitsize > 0 ? (uint32_t)cfg->bitsize : 255u;
    uint32_t c = cfg->c > 0 ? (uint32_t)cfg->c : pick_window(n, bits, factor, sizeof(F) == sizeof(fq_t));
    const msm_shape sh = make_msm_shape(n, c, bits, factor);
    uint32_t shift = c * sh.Wf;
    if (n) k_precompute_bases<F><<<grid_for(n, 64), 64, 0, st>>>(d_in, d_out, n, factor, shift, cfg->are_points_montgomery_form);
    if (!cfg->are_results_on_device) {
      e = cudaMemcpyAsync(out, d_out, sizeof(affine_t<F>) * (size_t)n * factor, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return map_cuda_error(e);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
  }
  if (!cfg->is_async) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return map_cuda_error(e);
  }
  return B381_SUCCESS;
}

template <class F>
static int combine_entry(const void* parts, int count, void* stream, bool on_device, void* result) {
  if (!parts || !result || count < 0) return B381_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  {
    Scratch sc(st);
    jacobian_t<F>* d_res = reinterpret_cast<jacobian_t<F>*>(result);
    if (!on_device && (e = sc.alloc(&d_res, 1)) != cudaSuccess) return map_cuda_error(e);
    k_msm_encode<F><<<1, 32, 0, st>>>(reinterpret_cast<const xyzz_t<F>*>(parts), count, false, d_res, 1u);
    if (!on_device) {
      e = cudaMemcpyAsync(result, d_res, sizeof(jacobian_t<F>), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return map_cuda_error(e);
    }
  }
  // a device result is ordered by the caller's stream (no host round trip per step when several MSMs are in flight);
  // a host result has to have landed when the call returns
  e = on_device ? cudaGetLastError() : cudaStreamSynchronize(st);
  return map_cuda_error(e);
}

}  // namespace b381

