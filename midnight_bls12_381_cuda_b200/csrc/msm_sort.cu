// Grouping of the MSM's (bucket, point) pairs and of its tasks: histogram / exclusive scan / scatter kernels written
// for this pipeline.  No library sort or scan is left on the MSM path (round 1 used cub::DeviceRadixSort and
// cub::DeviceScan here: 5.7 ms of a 2^24-point MSM).
//
// Replaces: the reference's thrust/CUB radix sort of 32-bit keys + histogram + inclusive scan
// (bls12-381/src/curve/msm_kernels.cu:224-256, :748-778).
#include "common.cuh"
#include "msm_core.cuh"
#include "msm_sort.cuh"

namespace b381 {

// ------------------------------------------------------------------ exclusive scan of uint32
// Tile of 4096 = 256 threads x 16 consecutive items: thread sums -> warp shuffles -> 8 warp totals -> exclusive prefixes
// written back; the tile's total goes to tile_sums[tile] (when there is more than one tile), which is scanned by the
// same kernel recursively and added back by k_scan_add.  Arrays here are bucket tables (<= a few MB), so three
// short launches beat a decoupled look-back single pass on latency.
constexpr int SCAN_TPB = 256, SCAN_IPT = 16, SCAN_TILE = SCAN_TPB * SCAN_IPT;

// base_dev / total_dev (single-tile launches only): a base value read from device memory and added to every prefix,
// and where base + total goes -- a chain of scans over consecutive pieces of one array needs no host round trip
static __global__ void __launch_bounds__(SCAN_TPB) k_scan_tile(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tile_sums,
                                                               uint32_t add, const uint32_t* base_dev, uint32_t* total_dev) {
  __shared__ uint32_t warp_tot[SCAN_TPB / 32];
  if (base_dev) add += *base_dev;                  // read before anything is written: base_dev may alias out[0]
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IPT;
  uint32_t v[SCAN_IPT];
  const bool wide = base + SCAN_IPT <= n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (wide) {
#pragma unroll
    for (int j = 0; j < SCAN_IPT / 4; j++) {
      const uint4 q = reinterpret_cast<const uint4*>(in + base)[j];
      v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_IPT; j++) v[j] = base + j < n ? in[base + j] : 0u;
  }
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < SCAN_IPT; j++) { const uint32_t x = v[j]; v[j] = sum; sum += x; }   // exclusive within the thread
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= (uint32_t)d) incl += y;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  uint32_t pre = add + incl - sum;                // exclusive prefix of this thread inside its warp (+ the caller's base)
  uint32_t tile_total = 0;
#pragma unroll
  for (int w = 0; w < SCAN_TPB / 32; w++) {
    const uint32_t t = warp_tot[w];
    if ((uint32_t)w < wid) pre += t;
    tile_total += t;
  }
  if (wide) {
#pragma unroll
    for (int j = 0; j < SCAN_IPT / 4; j++)
      reinterpret_cast<uint4*>(out + base)[j] = make_uint4(v[4 * j] + pre, v[4 * j + 1] + pre, v[4 * j + 2] + pre, v[4 * j + 3] + pre);
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_IPT; j++)
      if (base + j < n) out[base + j] = v[j] + pre;
  }
  if (tile_sums && threadIdx.x == 0) tile_sums[blockIdx.x] = tile_total;
  if (total_dev && threadIdx.x == 0) *total_dev = add + tile_total;
}

static __global__ void __launch_bounds__(256) k_scan_add(uint32_t* out, size_t n, const uint32_t* tile_offs, uint32_t* copy,
                                                         uint32_t base) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = out[i] + tile_offs[i / SCAN_TILE] + base;
  out[i] = v;
  if (copy) copy[i] = v;
}

static __global__ void __launch_bounds__(256) k_copy_u32(const uint32_t* in, uint32_t* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

cudaError_t exclusive_scan_u32(Scratch& sc, const uint32_t* in, uint32_t* out, size_t n, uint32_t* copy, int* launches,
                               uint32_t base, const uint32_t* base_dev, uint32_t* total_dev) {
  if (n == 0) return cudaSuccess;
  cudaStream_t st = sc.stream();
  const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles == 1) {
    k_scan_tile<<<1, SCAN_TPB, 0, st>>>(in, out, n, nullptr, base, base_dev, total_dev);
    if (launches) (*launches)++;
    if (copy) {
      k_copy_u32<<<grid_for(n, 256), 256, 0, st>>>(out, copy, n);
      if (launches) (*launches)++;
    }
    return cudaGetLastError();
  }
  uint32_t* sums;
  B381_CUDA_TRY(sc.alloc(&sums, tiles));
  k_scan_tile<<<(unsigned)tiles, SCAN_TPB, 0, st>>>(in, out, n, sums, 0u, nullptr, nullptr);
  if (launches) (*launches)++;
  // the device-side base and total travel with the scan of the tile totals: sums[i] = *base_dev + tiles before i
  B381_CUDA_TRY(exclusive_scan_u32(sc, sums, sums, tiles, nullptr, launches, 0u, base_dev, total_dev));
  k_scan_add<<<grid_for(n, 256), 256, 0, st>>>(out, n, sums, copy, base);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

// ------------------------------------------------------------------ (bucket, point) pairs: counting sort
static __global__ void __launch_bounds__(256) k_msm_hist(const fr_t* scalars, bool mont, msm_shape sh, uint32_t* hist,
                                                         uint32_t first, uint32_t end) {
  const uint32_t t = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < end) msm_hist_body(t, scalars, mont, sh, hist);
}

static __global__ void __launch_bounds__(256) k_msm_scatter(const fr_t* scalars, bool mont, msm_shape sh, uint32_t* cursor,
                                                            uint32_t* vals, uint32_t first, uint32_t end) {
  const uint32_t t = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < end) msm_scatter_body(t, scalars, mont, sh, cursor, vals);
}

// hist: msm_runs(sh) + 1 words of scratch; run_off: msm_runs(sh) + 1 words, the result.
// Resident scalars: one histogram kernel, one scan, one scatter kernel.
// Host scalars (host_src != nullptr; the plugin call with are_scalars_on_device = false): they cross PCIe in pieces on
// a copy stream (2^24 scalars = 512 MiB = 9.4 ms at 57 GB/s) and every piece is histogrammed as it lands; when the
// grouping is chunk-major a piece is a whole number of chunks, each chunk owns a fixed region of the sorted array,
// and the piece is scanned and scattered right away too, so the entire sort rides under the transfer.
bool msm_sort_is_streamed(const msm_shape& sh, const fr_t* host_src) {
  return host_src != nullptr && sh.nchunks > 1 && sh.batch == 1 && sh.n > msm_sort_piece(sh);
}
uint32_t msm_sort_piece(const msm_shape& sh) {
  uint32_t piece = 1u << 21;                            // 64 MiB of scalars per copy
  if (const char* e = getenv("B381_MSM_PIECE_LOG")) {   // tests: multi-piece runs at small sizes
    const int v = atoi(e);
    if (v >= 8 && v <= 26) piece = 1u << v;
  }
  if (sh.nchunks > 1) {
    // a piece is a whole number of chunks, and there are at least four pieces when there are four chunks: the last
    // piece's sort and forward pass are what the copy cannot hide
    const uint32_t chunk = 1u << sh.chunk_log;
    while (piece > chunk && (uint64_t)piece * 4 > sh.n) piece >>= 1;
    if (piece < chunk) piece = chunk;
  }
  return piece;
}

cudaError_t msm_sort_pairs(Scratch& sc, const fr_t* d_scalars, bool scalars_mont, const msm_shape& sh, const fr_t* host_src,
                           uint32_t* hist, uint32_t* run_off, uint32_t* vals, int* launches, const msm_piece_fn* after_piece) {
  cudaStream_t st = sc.stream();
  const uint32_t nt = sh.n * sh.batch;                  // scalars of the whole batch, [batch][n]
  const size_t nruns = msm_runs(sh);
  uint32_t* cursor;
  B381_CUDA_TRY(sc.alloc(&cursor, nruns + 1));
  // one spare slot: the scan of nruns + 1 entries leaves the pair total in run_off[nruns]
  B381_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (nruns + 1), st));
  const uint32_t piece = msm_sort_piece(sh);
  if (!host_src || nt <= piece) {
    if (host_src) B381_CUDA_TRY(cudaMemcpyAsync(const_cast<fr_t*>(d_scalars), host_src, sizeof(fr_t) * nt, cudaMemcpyHostToDevice, st));
    k_msm_hist<<<grid_for(nt, 256), 256, 0, st>>>(d_scalars, scalars_mont, sh, hist, 0u, nt);
    B381_CUDA_TRY(exclusive_scan_u32(sc, hist, run_off, nruns + 1, cursor, launches, 0u));
    k_msm_scatter<<<grid_for(nt, 256), 256, 0, st>>>(d_scalars, scalars_mont, sh, cursor, vals, 0u, nt);
    if (launches) *launches += 2;
    return cudaGetLastError();
  }
  cudaStream_t cp;
  cudaEvent_t ready, landed;
  B381_CUDA_TRY(cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking));
  cudaError_t e = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ready, st);             // d_scalars is a stream-ordered allocation of `st`
  if (e == cudaSuccess) e = cudaStreamWaitEvent(cp, ready, 0);
  if (e == cudaSuccess) cudaEventDestroy(ready);
  // chunks are self-contained: sort piece by piece (one MSM: in a batch the chunks of MSM b start at b * n, which
  // the fixed-size pieces of the flattened scalar array do not respect)
  const bool per_piece = sh.nchunks > 1 && sh.batch == 1;
  for (uint32_t first = 0; e == cudaSuccess && first < nt; first += piece) {
    const uint32_t end = nt - first < piece ? nt : first + piece;
    e = cudaMemcpyAsync(const_cast<fr_t*>(d_scalars) + first, host_src + first, sizeof(fr_t) * (end - first), cudaMemcpyHostToDevice, cp);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&landed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(landed, cp);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, landed, 0);
    if (e == cudaSuccess) cudaEventDestroy(landed);                 // released once the recorded work has completed
    if (e != cudaSuccess) break;
    k_msm_hist<<<grid_for(end - first, 256), 256, 0, st>>>(d_scalars, scalars_mont, sh, hist, first, end);
    if (launches) (*launches)++;
    if (per_piece) {
      const size_t r0 = (size_t)(first >> sh.chunk_log) * sh.nbuckets;
      const size_t r1 = end == nt ? nruns + 1 : (size_t)(end >> sh.chunk_log) * sh.nbuckets;   // last piece: + the total slot
      // one entry past the piece as well: run_off[r1] = where the NEXT piece starts (its counters are still zero), which
      // the slot walk of a streamed level 0 reads as the end of the piece's last run
      e = exclusive_scan_u32(sc, hist + r0, run_off + r0, (end == nt ? r1 : r1 + 1) - r0, cursor + r0, launches, first * sh.W);
      if (e != cudaSuccess) break;
      k_msm_scatter<<<grid_for(end - first, 256), 256, 0, st>>>(d_scalars, scalars_mont, sh, cursor, vals, first, end);
      if (launches) (*launches)++;
      // the piece's runs are final: the caller may start level 0 on them while the next piece is in flight
      if (after_piece && (e = (*after_piece)(r0, end == nt ? nruns : r1, end == nt)) != cudaSuccess) break;
    }
  }
  cudaStreamDestroy(cp);                                            // deferred by the runtime until its copies are done
  if (e != cudaSuccess) return e;
  if (!per_piece) {
    B381_CUDA_TRY(exclusive_scan_u32(sc, hist, run_off, nruns + 1, cursor, launches, 0u));
    k_msm_scatter<<<grid_for(nt, 256), 256, 0, st>>>(d_scalars, scalars_mont, sh, cursor, vals, 0u, nt);
    if (launches) (*launches)++;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ task visiting order: counting sort by length
// Keys 0..K (msm_task_key).  Most tasks share a handful of lengths, so the histogram is privatised in shared memory
// (same-address global atomics would serialise); K + 1 > TASK_BINS_SMEM only happens for huge mean bucket loads,
// where the keys are spread out and plain global atomics do.
constexpr uint32_t TASK_BINS_SMEM = 4096;

static __global__ void __launch_bounds__(256) k_msm_task_hist(const uint32_t* ntasks_dev, const uint2* tasks, uint32_t K,
                                                              uint32_t* hist, bool use_smem) {
  __shared__ uint32_t s_hist[TASK_BINS_SMEM];
  const uint32_t ntasks = *ntasks_dev, bins = K + 1;
  if ((uint64_t)blockIdx.x * blockDim.x >= ntasks) return;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (use_smem) {
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    if (t < ntasks) atomicAdd(&s_hist[msm_task_key(t, tasks, K)], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
      if (s_hist[b]) atomicAdd(hist + b, s_hist[b]);
  } else if (t < ntasks) {
    atomicAdd(hist + msm_task_key(t, tasks, K), 1u);
  }
}

static __global__ void __launch_bounds__(256) k_msm_task_scatter(const uint32_t* ntasks_dev, const uint2* tasks, uint32_t K,
                                                                 uint32_t* cursor, uint32_t* order, bool use_smem) {
  __shared__ uint32_t s_cnt[TASK_BINS_SMEM];      // per-bin count, then the CTA's base position in the bin
  const uint32_t ntasks = *ntasks_dev, bins = K + 1;
  if ((uint64_t)blockIdx.x * blockDim.x >= ntasks) return;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (use_smem) {
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    uint32_t key = 0, rank = 0;
    if (t < ntasks) {
      key = msm_task_key(t, tasks, K);
      rank = atomicAdd(&s_cnt[key], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
      if (s_cnt[b]) s_cnt[b] = atomicAdd(cursor + b, s_cnt[b]);
    __syncthreads();
    if (t < ntasks) order[s_cnt[key] + rank] = t;
  } else if (t < ntasks) {
    order[atomicAdd(cursor + msm_task_key(t, tasks, K), 1u)] = t;
  }
}

cudaError_t msm_task_order(Scratch& sc, size_t max_tasks, const uint32_t* ntasks_dev, const uint2* tasks, uint32_t K,
                           uint32_t* order, int* launches) {
  cudaStream_t st = sc.stream();
  const size_t bins = (size_t)K + 1;
  const bool use_smem = bins <= TASK_BINS_SMEM;
  uint32_t* hist;
  B381_CUDA_TRY(sc.alloc(&hist, bins));
  B381_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * bins, st));
  k_msm_task_hist<<<grid_for(max_tasks, 256), 256, 0, st>>>(ntasks_dev, tasks, K, hist, use_smem);
  B381_CUDA_TRY(exclusive_scan_u32(sc, hist, hist, bins, nullptr, launches, 0u));
  k_msm_task_scatter<<<grid_for(max_tasks, 256), 256, 0, st>>>(ntasks_dev, tasks, K, hist, order, use_smem);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

}  // namespace b381
