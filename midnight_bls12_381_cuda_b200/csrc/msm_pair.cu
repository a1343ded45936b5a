// Affine bucket pre-reduction: the three kernels of one level (bodies in msm_batch.cuh) for G1 (Fq) and
// G2 (Fq2), and their launcher.  Own translation unit so that tuning them does not recompile the rest of
// the MSM.
#include <cstdlib>

#include "common.cuh"
#include "msm_batch.cuh"

namespace b381 {

// thread t owns output slots [t*B, t*B + B); nt = threads in the grid = stride of the slot-major scratch
// chunk-major level 0 (msm_core.cuh): "buckets" are runs and dst_base / dstg carry where each slot's sum is written
struct pair_dst { const uint32_t* base; uint32_t set_slots, nchunks; uint32_t* slots; };
// streamed level 0 (msm_impl.cuh): the forward kernel is launched once per piece of runs while later pieces are still
// being copied and sorted.  A launch covers the threads whose 32 slots became complete with this piece: slot bounds
// [*lo_dev, *hi_dev) are device values (the scans that produce them run just before, on the same stream); runs beyond
// nb_search have no offsets yet.  hi_dev == nullptr: one launch over everything.
struct pair_range { const uint32_t* lo_dev; const uint32_t* hi_dev; uint32_t t0, nb_search, final; };

// CTAs per SM: 4 for G1 (<= 128 registers, no spill); an Fq2 forward pass spilled 408-480 bytes per thread at that cap
template <class F, int B, bool L0>
__global__ void __launch_bounds__(PR_TPB, (sizeof(F) > sizeof(fq_t) ? 2 : 4)) k_msm_pair_fwd(const uint32_t* in_off, const uint32_t* out_off,
                                                            uint32_t nbuckets, const uint32_t* svals,
                                                            const level_pts<F> pts, uint32_t nt, uint32_t* srcg,
                                                            F* preg, F* tot, const xrec_t<F>* xs, const pair_dst dst,
                                                            const pair_range rng) {
  const uint32_t t = rng.t0 + blockIdx.x * PR_TPB + threadIdx.x;
  uint32_t n_out, nb = nbuckets;
  if (rng.hi_dev) {
    const uint32_t hi = *rng.hi_dev;
    if (t >= nt || !pair_piece_owns(t, B, *rng.lo_dev, hi, rng.final != 0)) return;
    n_out = hi;
    nb = rng.nb_search;
  } else {
    n_out = out_off[nbuckets];
    if ((uint64_t)t * B >= n_out) return;
  }
  pair_walk<B>(t * B, n_out, in_off, out_off, nb, srcg + t, nt, dst.base, dst.set_slots, dst.nchunks,
               dst.base ? dst.slots + t : nullptr);
  tot[t] = pair_phase1<F, B, L0>(srcg + t, nt, svals, pts, preg + t, nt, xs);
}

// destinations of a chunk-major level 0 whose forward pass ran piece by piece (the bucket-major scan they come from
// needs every piece): the slot walk again, without the loads
template <int B>
__global__ void __launch_bounds__(PR_TPB) k_msm_pair_dst(const uint32_t* in_off, const uint32_t* out_off, uint32_t nbuckets,
                                                         uint32_t nt, const pair_dst dst) {
  const uint32_t n_out = out_off[nbuckets];
  const uint32_t t = blockIdx.x * PR_TPB + threadIdx.x;
  if ((uint64_t)t * B >= n_out) return;
  pair_walk<B>(t * B, n_out, in_off, out_off, nbuckets, nullptr, nt, dst.base, dst.set_slots, dst.nchunks, dst.slots + t);
}

template <class F>
__global__ void k_pack_x(const affine_t<F>* pts, size_t n, xrec_t<F>* xs) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) xs[i].x = pts[i].x;
}

// in-place inversion of the ceil(n_out / B) thread totals, up to M per thread
template <class F, int B, int M>
__global__ void __launch_bounds__(128) k_msm_invert_totals(const uint32_t* out_off, uint32_t nbuckets, F* tot) {
  const uint32_t n_out = out_off[nbuckets];
  const uint32_t count = (n_out + B - 1) / B;
  const uint32_t m = batch_invert_m(count, M);
  const uint32_t nthreads = (count + m - 1) / m;
  batch_invert_body<F, M>(blockIdx.x * blockDim.x + threadIdx.x, nthreads, count, m, tot);
}

template <class F, int B, bool L0, int MINB>
__global__ void __launch_bounds__(PR_TPB, MINB) k_msm_pair_bwd(const uint32_t* out_off, uint32_t nbuckets,
                                                               const uint32_t* svals, const level_pts<F> pts,
                                                               uint32_t nt, const uint32_t* srcg, const F* preg,
                                                               const F* tot, F* outx, F* outy, const uint32_t* dstg) {
  const uint32_t n_out = out_off[nbuckets];
  const uint32_t t = blockIdx.x * PR_TPB + threadIdx.x;
  if ((uint64_t)t * B >= n_out) return;
  if (dstg) pair_phase2<F, B, L0>(tot[t], srcg + t, nt, svals, pts, preg + t, nt, outx, outy, dstg + t);
  else pair_phase2<F, B, L0>(tot[t], srcg + t, nt, svals, pts, preg + t, nt, outx + (size_t)t * B, outy + (size_t)t * B);
}

// Live timing of the dominant kernels (level-0 forward and backward pass) for bench.py's roofline, recorded on
// the launching stream when B381_MSM_TIMING=1 and read back through b381_msm_last_level0_ms.
static thread_local cudaEvent_t g_l0_ev[4];
static thread_local bool g_l0_ev_init = false, g_l0_ev_valid = false;
static bool l0_timing_on() {
  static const bool on = [] { const char* e = getenv("B381_MSM_TIMING"); return e && e[0] == '1'; }();
  if (on && !g_l0_ev_init) {
    for (auto& e : g_l0_ev) cudaEventCreate(&e);
    g_l0_ev_init = true;
  }
  return on;
}

template <class F, bool L0>
static void launch_fwd(const uint32_t* in_off, const uint32_t* out_off, uint32_t nbuckets, const uint32_t* svals,
                       const level_pts<F> pts, size_t npts, unsigned g, uint32_t nt, uint32_t* srcg, F* preg, F* tot,
                       const pair_dst dst, const pair_range rng, cudaStream_t st) {
  constexpr int PB = pair_batch<F>::B;
  xrec_t<F>* xs = nullptr;
  // Optional (B381_XPACK=1): gather the level-0 x-coordinates from a packed, 64-byte-aligned copy.  Measured on
  // B200 at 2^24: forward pass 11.1 -> 10.3 ms (DRAM still fetches a 128-byte line per gather: 130 B/gather),
  // but building the copy costs 0.5 ms -- no net gain, so it is off by default.
  static const int xpack = [] { const char* e = getenv("B381_XPACK"); return e ? atoi(e) : 0; }();
  if (L0 && xpack && !rng.hi_dev && npts >= ((size_t)1 << 20) && cudaMallocAsync(&xs, npts * sizeof(xrec_t<F>), st) == cudaSuccess)
    k_pack_x<F><<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(pts.aos, npts, xs);
  else
    xs = nullptr;
  const bool timed = L0 && !rng.hi_dev && sizeof(F) == sizeof(fq_t) && l0_timing_on();
  if (timed) cudaEventRecord(g_l0_ev[0], st);
  k_msm_pair_fwd<F, PB, L0><<<g, PR_TPB, 0, st>>>(in_off, out_off, nbuckets, svals, pts, nt, srcg, preg, tot, xs, dst, rng);
  if (timed) cudaEventRecord(g_l0_ev[1], st);
  if (xs) cudaFreeAsync(xs, st);
}

template <class F, bool L0>
static void launch_bwd(const uint32_t* out_off, uint32_t nbuckets, const uint32_t* svals, const level_pts<F> pts,
                       unsigned g, uint32_t nt, uint32_t* srcg, F* preg, F* tot, F* outx, F* outy, const pair_dst dst,
                       bool timed, cudaStream_t st) {
  constexpr int PB = pair_batch<F>::B;
  // CTAs per SM of the backward kernel: G1 (126 registers at 4, no spill) measured best at 4 on B200
  // (profiles/r01b_msm_levels_sweep.txt); G2's Fq2 state needs the full register file: 254 registers, no spill at 2
  static const int minb = [] { const char* e = getenv("B381_BWD_MINB"); return e ? atoi(e) : 0; }();
  const int mb = minb ? minb : (sizeof(F) > sizeof(fq_t) ? 2 : 4);
  const uint32_t* dslots = dst.base ? dst.slots : nullptr;
  if (mb == 2) k_msm_pair_bwd<F, PB, L0, 2><<<g, PR_TPB, 0, st>>>(out_off, nbuckets, svals, pts, nt, srcg, preg, tot, outx, outy, dslots);
  else if (mb == 3) k_msm_pair_bwd<F, PB, L0, 3><<<g, PR_TPB, 0, st>>>(out_off, nbuckets, svals, pts, nt, srcg, preg, tot, outx, outy, dslots);
  else k_msm_pair_bwd<F, PB, L0, 4><<<g, PR_TPB, 0, st>>>(out_off, nbuckets, svals, pts, nt, srcg, preg, tot, outx, outy, dslots);
  if (timed) {
    cudaEventRecord(g_l0_ev[3], st);
    g_l0_ev_valid = true;
  }
}

template <class F, bool L0>
static void launch_invert_bwd(const uint32_t* out_off, uint32_t nbuckets, const uint32_t* svals, const level_pts<F> pts,
                              unsigned g, uint32_t nt, uint32_t* srcg, F* preg, F* tot, F* outx, F* outy, const pair_dst dst,
                              bool time_it, cudaStream_t st) {
  constexpr int PB = pair_batch<F>::B, PM = pair_batch<F>::M;
  const unsigned gi = (unsigned)((((size_t)nt + 3) / 4 + 127) / 128);   // enough for the smallest per-thread batch (4)
  const bool timed = time_it && L0 && sizeof(F) == sizeof(fq_t) && l0_timing_on();
  k_msm_invert_totals<F, PB, PM><<<gi, 128, 0, st>>>(out_off, nbuckets, tot);
  if (timed) cudaEventRecord(g_l0_ev[2], st);
  launch_bwd<F, L0>(out_off, nbuckets, svals, pts, g, nt, srcg, preg, tot, outx, outy, dst, timed, st);
}

// A level pays once there are enough pairs to fill the GPU (each level has ~0.2 ms of fixed latency: three
// launches and one thread-serial inversion) and buckets are still long.  Measured on B200 (gpurun sweep of
// B381_MSM_LEVELS, profiles/r01c_msm_levels_sweep.txt): 2^18 points -> 1 level, 2^20 -> 3, 2^21 -> 4, 2^22 -> 5,
// 2^24 -> 6 or 7 (equal), i.e. levels down to 2^21 pair sums.
int msm_pair_levels(double avg, size_t total) {
  int levels = 0;
  while (levels < 8 && avg >= 8.0 && total / 2 >= ((size_t)1 << 21)) { levels++; avg *= 0.5; total /= 2; }
  return levels;
}

template <class F>
void launch_pair_level(bool level0, const uint32_t* in_off, const uint32_t* out_off, uint32_t nbuckets,
                       const uint32_t* svals, const level_pts<F> pts, size_t npts, unsigned grid, uint32_t* srcg, F* preg,
                       F* tot, F* outx, F* outy, cudaStream_t st, const uint32_t* dst_base, uint32_t dst_set_slots,
                       uint32_t dst_nchunks, uint32_t* dst_slots) {
  const pair_dst dst{dst_base, dst_set_slots, dst_nchunks, dst_slots};
  const pair_range all{nullptr, nullptr, 0u, 0u, 0u};
  const uint32_t nt = grid * PR_TPB;
  if (level0) {
    launch_fwd<F, true>(in_off, out_off, nbuckets, svals, pts, npts, grid, nt, srcg, preg, tot, dst, all, st);
    launch_invert_bwd<F, true>(out_off, nbuckets, svals, pts, grid, nt, srcg, preg, tot, outx, outy, dst, true, st);
  } else {
    launch_fwd<F, false>(in_off, out_off, nbuckets, nullptr, pts, npts, grid, nt, srcg, preg, tot, dst, all, st);
    launch_invert_bwd<F, false>(out_off, nbuckets, nullptr, pts, grid, nt, srcg, preg, tot, outx, outy, dst, true, st);
  }
}

// streamed chunk-major level 0, forward part for the runs [.., nb_search) of one piece (see pair_range): threads
// t0 .. t0 + piece_grid * PR_TPB of a scratch layout with stride `grid` * PR_TPB
template <class F>
void launch_pair_fwd_piece(const uint32_t* in_off, const uint32_t* out_off, const uint32_t* svals, const level_pts<F> pts,
                           unsigned grid, unsigned piece_grid, uint32_t t0, uint32_t nb_search, bool final,
                           const uint32_t* lo_dev, const uint32_t* hi_dev, uint32_t* srcg, F* preg, F* tot, cudaStream_t st) {
  const pair_dst none{nullptr, 0u, 0u, nullptr};
  const pair_range rng{lo_dev, hi_dev, t0, nb_search, final ? 1u : 0u};
  launch_fwd<F, true>(in_off, out_off, nb_search, svals, pts, 0, piece_grid, grid * PR_TPB, srcg, preg, tot, none, rng, st);
}

// ... and its second half once every piece is in: destinations, batched inversion, backward pass
template <class F>
void launch_pair_finish_streamed(const uint32_t* in_off, const uint32_t* out_off, uint32_t nbuckets, const uint32_t* svals,
                                 const level_pts<F> pts, unsigned grid, uint32_t* srcg, F* preg, F* tot, F* outx, F* outy,
                                 cudaStream_t st, const uint32_t* dst_base, uint32_t dst_set_slots, uint32_t dst_nchunks,
                                 uint32_t* dst_slots) {
  const pair_dst dst{dst_base, dst_set_slots, dst_nchunks, dst_slots};
  const uint32_t nt = grid * PR_TPB;
  // The destination walk and the batched inversion are independent (only the backward pass needs both), and the
  // inversion is latency-bound at 11-20 % occupancy: the walk (0.5 ms at 2^24) runs beside it on a side stream.
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool forked = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&join, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventRecord(fork, st) == cudaSuccess && cudaStreamWaitEvent(side, fork, 0) == cudaSuccess;
  k_msm_pair_dst<pair_batch<F>::B><<<grid, PR_TPB, 0, forked ? side : st>>>(in_off, out_off, nbuckets, nt, dst);
  if (forked) forked = cudaEventRecord(join, side) == cudaSuccess;
  constexpr int PB = pair_batch<F>::B, PM = pair_batch<F>::M;
  const unsigned gi = (unsigned)((((size_t)nt + 3) / 4 + 127) / 128);
  k_msm_invert_totals<F, PB, PM><<<gi, 128, 0, st>>>(out_off, nbuckets, tot);
  if (forked) cudaStreamWaitEvent(st, join, 0);
  else if (side) cudaStreamSynchronize(side);    // could not record the join: wait on the host (the walk may be on `side`)
  launch_bwd<F, true>(out_off, nbuckets, svals, pts, grid, nt, srcg, preg, tot, outx, outy, dst, false, st);
  if (fork) cudaEventDestroy(fork);
  if (join) cudaEventDestroy(join);
  if (side) cudaStreamDestroy(side);             // deferred by the runtime until the walk is done
}

#define B381_INSTANTIATE_PAIR(F)                                                                                          \
  template void launch_pair_level<F>(bool, const uint32_t*, const uint32_t*, uint32_t, const uint32_t*, const level_pts<F>, \
                                     size_t, unsigned, uint32_t*, F*, F*, F*, F*, cudaStream_t, const uint32_t*, uint32_t,  \
                                     uint32_t, uint32_t*);                                                                  \
  template void launch_pair_fwd_piece<F>(const uint32_t*, const uint32_t*, const uint32_t*, const level_pts<F>, unsigned,   \
                                         unsigned, uint32_t, uint32_t, bool, const uint32_t*, const uint32_t*, uint32_t*,   \
                                         F*, F*, cudaStream_t);                                                             \
  template void launch_pair_finish_streamed<F>(const uint32_t*, const uint32_t*, uint32_t, const uint32_t*,                 \
                                               const level_pts<F>, unsigned, uint32_t*, F*, F*, F*, F*, cudaStream_t,        \
                                               const uint32_t*, uint32_t, uint32_t, uint32_t*);
B381_INSTANTIATE_PAIR(fq_t)
B381_INSTANTIATE_PAIR(fq2_t)

}  // namespace b381

extern "C" int b381_msm_last_level0_ms(float* fwd_ms, float* bwd_ms) {
  using namespace b381;
  if (!g_l0_ev_valid || !fwd_ms || !bwd_ms) return 0;
  if (cudaEventSynchronize(g_l0_ev[3]) != cudaSuccess) return 0;
  if (cudaEventElapsedTime(fwd_ms, g_l0_ev[0], g_l0_ev[1]) != cudaSuccess) return 0;
  if (cudaEventElapsedTime(bwd_ms, g_l0_ev[2], g_l0_ev[3]) != cudaSuccess) return 0;
  return 1;
}
