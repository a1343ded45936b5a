// Fr NTT / iNTT for sm_100a: domain (twiddle) management, pass planner, kernels, C ABI.
// Butterfly bodies are in ntt_core.cuh.
//
// Replaces: init_domain_cuda_impl / release_domain_cuda_impl / ntt_cuda_impl / coset_ntt_cuda_impl
// (bls12-381/src/field/ntt_kernels.cu:1607-1679, :1823-1848, :968-1133, :1155-1306) and the
// wrappers in bls12-381/src/backend/icicle_field_api.cu:97-131.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "common.cuh"
#include "ntt_core.cuh"

namespace b381 {

constexpr uint32_t kTile = 1u << kNttTileLog;
constexpr uint32_t kMaxDomainLog = 27;            // two tables of 2^27 - 1 twiddles = 8 GiB; larger roots are refused

// ------------------------------------------------------------------ kernels
// One pass = one kernel: tile load, (S + 1) / 2 register-blocked steps with a CTA barrier after each, tile store.
// S (stages of the pass) and the tile geometry are template parameters, so every slot / twiddle index in the steps is
// shifts and masks by constants:
//   LO0 = true : last pass (lo = 0): contiguous 2^TL-element tile = 2^(TL-S) blocks of 2^S, no adjacent-column bits;
//   LO0 = false: upper pass: 2^S rows x 2^(TL-S) adjacent columns.
// NT threads per CTA, MINB CTAs per SM (register cap via __launch_bounds__).
template <int S, bool LO0, int NT, int MINB, int UNR = 1, bool PF = false>
__global__ void __launch_bounds__(NT, MINB) k_ntt(const __grid_constant__ ntt_pass_params p, const fr_t* in, fr_t* out) {
  extern __shared__ uint4 smem[];
  const ntt_tile t{smem, smem + kTile};
  const uint64_t tile_id = p.tile_rot ? (blockIdx.x + p.tile_rot) % gridDim.x : blockIdx.x;
  const ntt_tile_ctx c = ntt_tile_begin(p, tile_id, (uint32_t)S, LO0 ? 0u : (uint32_t)(kNttTileLog - S),
                                        LO0 ? (uint32_t)(kNttTileLog - S) : 0u, LO0 ? 0u : p.lo);
#pragma unroll 4
  for (uint32_t pos = threadIdx.x; pos < kTile; pos += NT) ntt_tile_load(p, c, pos, in, t);
  cp_async_wait_all();
  __syncthreads();
  constexpr uint32_t STEPS = (S + 1) / 2;
#pragma unroll
  for (uint32_t step = 0; step < STEPS; step++) {
    uint32_t rl, s;
    ntt_pass_step(S, step, &rl, &s);
    if (rl == 1) {
#pragma unroll 1
      for (uint32_t q = threadIdx.x; q < kTile / 2; q += NT) ntt_step_r2(p, c, q, s, t);
    } else if (LO0 && s == 0) {
#pragma unroll UNR
      for (uint32_t q = threadIdx.x; q < kTile / 4; q += NT) ntt_step_r4<true>(p, c, q, s, t);
    } else {
#pragma unroll UNR
      for (uint32_t q = threadIdx.x; q < kTile / 4; q += NT) {
        if (PF) {            // twiddles of the thread's next group: same step, or the first group of the next step
          if (q + NT < kTile / 4) ntt_prefetch_r4(p, c, q + NT, s);
          else if (s >= 2 && !(LO0 && s == 2)) ntt_prefetch_r4(p, c, threadIdx.x, s - 2);
        }
        ntt_step_r4<false>(p, c, q, s, t);
      }
    }
    __syncthreads();
  }
#pragma unroll 2
  for (uint32_t pos = threadIdx.x; pos < kTile; pos += NT) ntt_tile_store(p, c, pos, out, t);
}

// Same pass with every shape parameter at run time: tiles smaller than 2^TL (distributed column passes over few local
// columns) and the B381_NTT_GENERIC=1 cross-check of the specialised kernels.
__global__ void __launch_bounds__(256, 2) k_ntt_generic(const __grid_constant__ ntt_pass_params p, const fr_t* in, fr_t* out) {
  extern __shared__ uint4 smem[];
  const uint32_t tile = 1u << (p.S + p.g + p.x);
  const ntt_tile t{smem, smem + tile};
  const uint64_t tile_id = p.tile_rot ? (blockIdx.x + p.tile_rot) % gridDim.x : blockIdx.x;
  const ntt_tile_ctx c = ntt_tile_begin(p, tile_id, p.S, p.g, p.x, p.lo);
  for (uint32_t pos = threadIdx.x; pos < tile; pos += blockDim.x) ntt_tile_load(p, c, pos, in, t);
  cp_async_wait_all();
  __syncthreads();
  const uint32_t steps = ntt_pass_steps(p.S);
  const bool last_is_unit = p.lo == 0 && p.dist_shift == 0;
  for (uint32_t step = 0; step < steps; step++) {
    uint32_t rl, s;
    ntt_pass_step(p.S, step, &rl, &s);
    if (rl == 1) {
      for (uint32_t q = threadIdx.x; q < tile / 2; q += blockDim.x) ntt_step_r2(p, c, q, s, t);
    } else if (last_is_unit && s == 0) {
      for (uint32_t q = threadIdx.x; q < tile / 4; q += blockDim.x) ntt_step_r4<true>(p, c, q, s, t);
    } else {
      for (uint32_t q = threadIdx.x; q < tile / 4; q += blockDim.x) ntt_step_r4<false>(p, c, q, s, t);
    }
    __syncthreads();
  }
  for (uint32_t pos = threadIdx.x; pos < tile; pos += blockDim.x) ntt_tile_store(p, c, pos, out, t);
}

// launch shape of the specialised kernels: 256 threads, 3 CTAs per SM (80 registers); B381_NTT_SHAPE=1 = 2 CTAs per SM
// (up to 128 registers) for A/B runs.  Measured on B200 at 2^24 (profiles/r02_ntt_variants.txt): 3.60 ms vs 3.69 ms;
// 128 or 512 threads, two groups per thread interleaved, and L1 prefetch of the next group's twiddles were all slower
// or equal and are not built.
struct ntt_shape { int nt, minb; };
static int ntt_shape_id() {
  static const int v = [] { const char* e = getenv("B381_NTT_SHAPE"); int x = e ? atoi(e) : 0; return x < 0 || x > 1 ? 0 : x; }();
  return v;
}
static bool ntt_force_generic() {
  static const bool v = [] { const char* e = getenv("B381_NTT_GENERIC"); return e && e[0] == '1'; }();
  return v;
}

template <int S, bool LO0, int NT, int MINB, int UNR = 1, bool PF = false>
static void launch_k(const ntt_pass_params& p, const fr_t* in, fr_t* out, unsigned tiles, cudaStream_t st) {
  constexpr size_t smem = 2 * sizeof(uint4) * kTile;
  static bool attr_done = false;       // per instantiation
  if (!attr_done) {
    cudaFuncSetAttribute(k_ntt<S, LO0, NT, MINB, UNR, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  k_ntt<S, LO0, NT, MINB, UNR, PF><<<tiles, NT, smem, st>>>(p, in, out);
}
template <int S, bool LO0>
static void launch_shape(const ntt_pass_params& p, const fr_t* in, fr_t* out, unsigned tiles, cudaStream_t st) {
  switch (ntt_shape_id()) {
    case 1: launch_k<S, LO0, 256, 2>(p, in, out, tiles, st); break;
    default: launch_k<S, LO0, 256, 3>(p, in, out, tiles, st); break;
  }
}
template <bool LO0, int S>
static void launch_s(uint32_t Sr, const ntt_pass_params& p, const fr_t* in, fr_t* out, unsigned tiles, cudaStream_t st) {
  if constexpr (S == 0) {
    (void)Sr; (void)p; (void)in; (void)out; (void)tiles; (void)st;
  } else {
    if (Sr == (uint32_t)S) launch_shape<S, LO0>(p, in, out, tiles, st);
    else launch_s<LO0, S - 1>(Sr, p, in, out, tiles, st);
  }
}

static void launch_ntt_pass(const ntt_pass_params& p, const fr_t* in, fr_t* out, unsigned tiles, cudaStream_t st) {
  const uint32_t tile_log = p.S + p.g + p.x;
  const bool full = tile_log == kNttTileLog && !ntt_force_generic();
  if (full && p.lo == 0 && p.dist_shift == 0 && p.g == 0 && p.S >= 1 && p.S <= kNttTileLog) {
    launch_s<true, (int)kNttTileLog>(p.S, p, in, out, tiles, st);
  } else if (full && p.lo > 0 && p.x == 0 && p.S >= 1 && p.S <= kNttTileLog - 1) {
    launch_s<false, (int)kNttTileLog - 1>(p.S, p, in, out, tiles, st);
  } else {
    const size_t smem = (size_t)2 * sizeof(uint4) << tile_log;
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(k_ntt_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(uint4) * kTile));
      attr_done = true;
    }
    k_ntt_generic<<<tiles, 256, smem, st>>>(p, in, out);
  }
}

// out[j] = base * g^j, j < count, 64 consecutive powers per thread
__global__ void k_fr_powers(fr_t g, fr_t base, uint64_t count, fr_t* out) {
  uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_powers_chunk(chunk, 64, count, g, base, out);
}

// stage-major table from the top-level powers: T[2^k - 1 + j] = top[j << (K-1-k)]
__global__ void k_twiddle_subsample(const fr_t* top, uint32_t K, fr_t* table) {
  uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. 2^(K-1)-2 : all levels below the top
  uint64_t lim = (1ull << (K - 1)) - 1;
  if (idx >= lim) return;
  uint32_t k = 63 - __clzll(idx + 1);
  uint64_t j = idx + 1 - (1ull << k);
  table[idx] = top[j << (K - 1 - k)];
}

// inverse stage-major table from the forward one (ntt_core.cuh ntt_inverse_twiddle)
__global__ void k_twiddle_inverse(const fr_t* fwd, uint64_t entries, fr_t* inv_table) {
  uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < entries) ntt_inverse_twiddle(idx, fwd, inv_table);
}

// single-thread setup: decide whether `root` is standard or Montgomery, find its order, n^-1 table
struct domain_probe { int log_order; int was_montgomery; fr_t root_mont; fr_t ninv[33]; };
__global__ void k_domain_probe(fr_t root, domain_probe* out) {
  if (blockIdx.x || threadIdx.x) return;
  fr_t cand[2] = {to_mont(root), root};       // reading 0: bytes are standard form; 1: already Montgomery
  fr_t minus1 = neg(one<fr_t>());
  out->log_order = -1;
  for (int r = 0; r < 2; r++) {
    fr_t x = cand[r];
    if (is_zero(x)) continue;
    // order 2^k  <=>  x^(2^(k-1)) = -1  (k >= 1);  order 1 <=> x = 1
    if (eq(x, one<fr_t>())) { out->log_order = 0; out->was_montgomery = r; out->root_mont = x; break; }
    int found = -1;
    for (int k = 1; k <= 32; k++) {
      if (eq(x, minus1)) { found = k; break; }
      x = sqr(x);
    }
    if (found >= 0) { out->log_order = found; out->was_montgomery = r; out->root_mont = cand[r]; break; }
  }
  fr_t two = add(one<fr_t>(), one<fr_t>());
  fr_t half = inv(two);
  fr_t acc = one<fr_t>();
  for (int k = 0; k <= 32; k++) { out->ninv[k] = acc; acc = mul(acc, half); }
}

// ------------------------------------------------------------------ domain
struct CosetTable { fr_t gen; uint32_t n; bool inverse; fr_t* dev; };

struct Domain {
  bool ready = false;
  int device = -1;
  uint32_t K = 0;            // log2 of the root's order
  fr_t root_mont;            // primitive 2^K-th root, Montgomery
  fr_t* table = nullptr;     // 2^K - 1 twiddles, stage-major; followed by the same for omega^-1
  fr_t* inv_table = nullptr;
  fr_t ninv[33];             // 2^-k, Montgomery
  std::vector<CosetTable> cosets;
};
// One domain PER DEVICE (ICICLE's set_device model: a single process may drive several GPUs; twiddles built on
// device 0 are not addressable from device 1).  Every entry point works on the calling thread's current device.
static std::map<int, Domain> g_doms;
static std::mutex g_dom_mu;
static thread_local int g_ntt_last_info[4] = {0, 0, 0, 0};   // passes, log n, batch, tiles of the last pass

// caller holds g_dom_mu
static Domain& cur_domain() {
  int dev = 0;
  cudaGetDevice(&dev);
  Domain& d = g_doms[dev];
  d.device = dev;
  return d;
}

static void release_locked() {
  Domain& d = cur_domain();
  if (d.table) cudaFree(d.table);
  for (auto& c : d.cosets) cudaFree(c.dev);
  const int dev = d.device;
  d = Domain();
  d.device = dev;
}

static int init_domain(const fr_t& root, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_dom_mu);
  Domain& g_dom = cur_domain();
  if (g_dom.ready) return B381_SUCCESS;   // ICICLE semantics: repeated init is a no-op until release
  domain_probe* d_probe;
  domain_probe h;
  if (cudaMalloc(&d_probe, sizeof(domain_probe)) != cudaSuccess) return B381_ALLOCATION_FAILED;
  k_domain_probe<<<1, 1, 0, st>>>(root, d_probe);
  cudaError_t e = cudaMemcpyAsync(&h, d_probe, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_probe);
  if (e != cudaSuccess) return map_cuda_error(e);
  if (h.log_order < 1 || h.log_order > (int)kMaxDomainLog) return B381_INVALID_ARGUMENT;
  const uint32_t K = (uint32_t)h.log_order;
  fr_t* table;
  const uint64_t entries = (1ull << K) - 1;
  if (cudaMalloc(&table, 2 * entries * sizeof(fr_t)) != cudaSuccess) return B381_OUT_OF_MEMORY;
  // top level: T[2^(K-1)-1 + j] = root^j, j < 2^(K-1)
  fr_t* top = table + ((1ull << (K - 1)) - 1);
  const uint64_t half = 1ull << (K - 1);
  fr_t one_m = {FR_ONE_INIT};
  k_fr_powers<<<grid_for((half + 63) / 64, 128), 128, 0, st>>>(h.root_mont, one_m, half, top);
  if (K > 1) k_twiddle_subsample<<<grid_for(half - 1, 256), 256, 0, st>>>(top, K, table);
  k_twiddle_inverse<<<grid_for(entries, 256), 256, 0, st>>>(table, entries, table + entries);
  e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { cudaFree(table); return map_cuda_error(e); }
  g_dom.K = K;
  g_dom.root_mont = h.root_mont;
  g_dom.table = table;
  g_dom.inv_table = table + entries;
  memcpy(g_dom.ninv, h.ninv, sizeof(h.ninv));
  g_dom.ready = true;
  return B381_SUCCESS;
}

static bool fr_host_eq(const fr_t& a, const fr_t& b) { return memcmp(&a, &b, sizeof(fr_t)) == 0; }
static bool fr_host_is_zero(const fr_t& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }

__global__ void k_fr_inv1(fr_t x, fr_t* out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = inv(x);
}

// Coset power tables, cached per (generator, log size, direction) -- the reference caches only the
// first generator it sees (ntt_kernels.cu:1701-1705).  forward: g^i ; inverse: g^-k * 2^-n.
// Caller holds g_dom_mu.
static int coset_table(Domain& g_dom, const fr_t& gen, uint32_t n, bool inverse, cudaStream_t st, const fr_t** out) {
  for (auto& c : g_dom.cosets)
    if (c.n == n && c.inverse == inverse && fr_host_eq(c.gen, gen)) { *out = c.dev; return B381_SUCCESS; }
  if (g_dom.cosets.size() >= 8) {          // small FIFO cache; eviction is rare, so a full sync is fine
    cudaDeviceSynchronize();
    cudaFree(g_dom.cosets.front().dev);
    g_dom.cosets.erase(g_dom.cosets.begin());
  }
  fr_t* dev;
  const uint64_t N = 1ull << n;
  if (cudaMalloc(&dev, N * sizeof(fr_t)) != cudaSuccess) return B381_OUT_OF_MEMORY;
  fr_t g = gen, base = {FR_ONE_INIT};
  if (inverse) {
    base = g_dom.ninv[n];
    fr_t* d_g;
    if (cudaMalloc(&d_g, sizeof(fr_t)) != cudaSuccess) { cudaFree(dev); return B381_OUT_OF_MEMORY; }
    k_fr_inv1<<<1, 1, 0, st>>>(gen, d_g);
    cudaError_t e = cudaMemcpyAsync(&g, d_g, sizeof(fr_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_g);
    if (e != cudaSuccess) { cudaFree(dev); return map_cuda_error(e); }
  }
  k_fr_powers<<<grid_for((N + 63) / 64, 128), 128, 0, st>>>(g, base, N, dev);
  // The table is published to every stream through the cache, so it must be COMPLETE before it becomes visible: a
  // cache hit on another stream has no dependency on `st`.  Builds are rare (once per generator and size).
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { cudaFree(dev); return B381_UNKNOWN_ERROR; }
  g_dom.cosets.push_back(CosetTable{gen, n, inverse, dev});
  *out = dev;
  return B381_SUCCESS;
}

// ------------------------------------------------------------------ launcher
static int ntt_run(const fr_t* input, int size, int dir, const b381_ntt_config* cfg, fr_t* output) {
  if (!cfg) return B381_INVALID_POINTER;
  if (size < 0) return B381_INVALID_ARGUMENT;
  if (size == 0) return B381_SUCCESS;
  if (!input || !output) return B381_INVALID_POINTER;
  if (size & (size - 1)) return B381_INVALID_ARGUMENT;       // power of two (ntt_kernels.cu:717-722)
  uint32_t n = 0;
  while ((1u << n) < (uint32_t)size) n++;
  const uint32_t batch = cfg->batch_size > 0 ? (uint32_t)cfg->batch_size : 1u;
  const uint64_t total = (uint64_t)batch << n;
  const bool inverse = dir == B381_NTT_INVERSE;
  cudaStream_t st = (cudaStream_t)cfg->stream;

  std::unique_lock<std::mutex> lk(g_dom_mu);
  Domain& g_dom = cur_domain();
  if (!g_dom.ready || n > g_dom.K) return B381_INVALID_ARGUMENT;   // domain missing / too small (:724-731)
  const fr_t* tw = inverse ? g_dom.inv_table : g_dom.table;
  fr_t ninv = g_dom.ninv[n];
  const fr_t one_m = {FR_ONE_INIT};
  const fr_t* gen = reinterpret_cast<const fr_t*>(&cfg->coset_gen);
  bool has_coset = !fr_host_eq(*gen, one_m) && !fr_host_is_zero(*gen);
  const fr_t* ctab = nullptr;
  if (has_coset) {
    int rc = coset_table(g_dom, *gen, n, inverse, st, &ctab);
    if (rc != B381_SUCCESS) return rc;
  }

  bool perm_in = false, perm_out = false;
  switch (cfg->ordering) {
    case B381_kNN: perm_out = true; break;
    case B381_kNR: case B381_kNM: break;
    case B381_kRN: case B381_kMN: perm_in = true; perm_out = true; break;
    case B381_kRR: perm_in = true; break;
    default: return B381_INVALID_ARGUMENT;
  }

  cudaError_t e;
  TraceRange trace(inverse ? "b381_ntt inverse" : "b381_ntt forward", (long long)total);
  {
    Scratch sc(st);
    const fr_t* d_in;
    if ((e = stage_in(sc, input, total, cfg->are_inputs_on_device, &d_in)) != cudaSuccess) return map_cuda_error(e);
    fr_t* d_out = output;
    if (!cfg->are_outputs_on_device && (e = sc.alloc(&d_out, total)) != cudaSuccess) return map_cuda_error(e);

    ntt_pass_plan plan[8];
    size_t P = 0;
    if (const char* e = getenv("B381_NTT_PLAN")) {     // e.g. "10,10": stage counts, highest first (A/B runs only)
      uint32_t S[8], cnt = 0;
      for (const char* q = e; *q && cnt < 8;) {
        S[cnt++] = (uint32_t)strtoul(q, const_cast<char**>(&q), 10);
        if (*q == ',') q++;
      }
      P = (size_t)ntt_plan_from_list(n, S, cnt, plan);
    }
    if (P == 0) P = (size_t)ntt_plan_passes(n, plan);
    const bool inplace = (d_in == d_out);
    fr_t* work = d_out;
    if (P >= 2 && (perm_in || perm_out) && (inplace || perm_out)) {
      if ((e = sc.alloc(&work, total)) != cudaSuccess) return map_cuda_error(e);
    }
    for (size_t i = 0; i < P; i++) {
      ntt_pass_params p;
      memset(&p, 0, sizeof(p));
      p.n = n; p.lo = plan[i].lo; p.S = plan[i].S; p.g = plan[i].g; p.x = plan[i].x;
      p.total = total;
      if (cfg->columns_batch) { p.estride = batch; p.bstride = 1; }
      else { p.estride = 1; p.bstride = 1ull << n; }
      p.twiddles = tw;
      const bool first = (i == 0), last = (i + 1 == P);
      p.perm_in = first && perm_in;
      p.perm_out = last && perm_out;
      if (first && has_coset && !inverse) p.pre_scale = ctab;
      if (last && inverse) {
        if (has_coset) p.post_scale = ctab;          // g^-k * 2^-n
        else { p.post_const = ninv; p.has_post_const = 1; }
      }
      const fr_t* src = first ? d_in : work;
      fr_t* dst = last ? d_out : work;
      const uint32_t tile_log = p.S + p.g + p.x;
      const uint64_t tiles = (total + (1ull << tile_log) - 1) >> tile_log;
      launch_ntt_pass(p, src, dst, (unsigned)tiles, st);
    }
    g_ntt_last_info[0] = (int)P; g_ntt_last_info[1] = (int)n; g_ntt_last_info[2] = (int)batch;
    lk.unlock();   // tables stay valid: release_domain synchronises the device before freeing
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->are_outputs_on_device) {
      e = cudaMemcpyAsync(output, d_out, total * sizeof(fr_t), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return map_cuda_error(e);
    }
  }
  if (!cfg->is_async) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return map_cuda_error(e);
  }
  return B381_SUCCESS;
}

// Step 1 of the distributed (four-step) NTT: the top `a` stages of a 2^log_n transform on this
// GPU's column block (see ntt_pass_params::dist_*), in place, device memory.
static int ntt_dist_columns(fr_t* data, uint32_t log_n, uint32_t log_gpus, uint32_t rank, uint32_t a, int dir,
                            cudaStream_t st, void* const* peer_rows = nullptr) {
  if (!data) return B381_INVALID_POINTER;
  if (log_gpus == 0 || a == 0 || a + log_gpus > log_n || rank >= (1u << log_gpus)) return B381_INVALID_ARGUMENT;
  const uint32_t lo = log_n - a;            // global bit where the column index ends
  if (lo < log_gpus + 2) return B381_INVALID_ARGUMENT;   // keep >= 4 adjacent elements per tile row
  if (peer_rows && log_gpus > 3) return B381_INVALID_ARGUMENT;   // ntt_pass_params::peer_out holds 8 GPUs
  std::unique_lock<std::mutex> lk(g_dom_mu);
  Domain& g_dom = cur_domain();
  if (!g_dom.ready || log_n > g_dom.K) return B381_INVALID_ARGUMENT;
  const fr_t* tw = dir == B381_NTT_INVERSE ? g_dom.inv_table : g_dom.table;
  ntt_pass_params passes[8];
  const int np = ntt_dist_passes(log_n, log_gpus, rank, a, tw, passes);
  for (int i = 0; i < np; i++) {
    ntt_pass_params& p = passes[i];
    if (peer_rows && i + 1 == np) {
      p.peer_on = 1;
      p.peer_logR = a - log_gpus;
      for (uint32_t r = 0; r < (1u << log_gpus); r++) {
        if (!peer_rows[r]) return B381_INVALID_POINTER;
        p.peer_out[r] = (fr_t*)peer_rows[r];
      }
    }
    const uint32_t tile_log = p.S + p.g;
    const uint64_t tiles = p.total >> tile_log;
    if (p.peer_on) p.tile_rot = (uint32_t)((tiles >> log_gpus) * rank);   // tiles are ordered by destination GPU
    launch_ntt_pass(p, data, data, (unsigned)tiles, st);
  }
  lk.unlock();
  return map_cuda_error(cudaGetLastError());
}

}  // namespace b381
using namespace b381;

extern "C" {
int b381_ntt_dist_columns(b381_fr* data_device, int log_n, int log_gpus, int rank, int upper_stages, int dir, void* stream) {
  if (log_n < 0 || log_gpus < 0 || rank < 0 || upper_stages < 0) return B381_INVALID_ARGUMENT;
  return ntt_dist_columns((fr_t*)data_device, (uint32_t)log_n, (uint32_t)log_gpus, (uint32_t)rank, (uint32_t)upper_stages, dir,
                          (cudaStream_t)stream);
}
int b381_ntt_dist_columns_p2p(b381_fr* data_device, int log_n, int log_gpus, int rank, int upper_stages, int dir,
                              void* const* peer_rows, void* stream) {
  if (log_n < 0 || log_gpus < 0 || rank < 0 || upper_stages < 0) return B381_INVALID_ARGUMENT;
  if (!peer_rows) return B381_INVALID_POINTER;
  return ntt_dist_columns((fr_t*)data_device, (uint32_t)log_n, (uint32_t)log_gpus, (uint32_t)rank, (uint32_t)upper_stages, dir,
                          (cudaStream_t)stream, peer_rows);
}
int b381_ntt_init_domain(const b381_fr* root, const b381_ntt_init_domain_config* cfg) {
  if (!root) return B381_INVALID_POINTER;
  cudaStream_t st = cfg ? (cudaStream_t)cfg->stream : 0;
  return init_domain(*reinterpret_cast<const fr_t*>(root), st);
}
int b381_ntt_release_domain(void) {
  std::lock_guard<std::mutex> lk(g_dom_mu);
  cudaDeviceSynchronize();
  release_locked();
  return B381_SUCCESS;
}
int b381_ntt(const b381_fr* in, int size, int dir, const b381_ntt_config* cfg, b381_fr* out) {
  return ntt_run((const fr_t*)in, size, dir, cfg, (fr_t*)out);
}
int b381_ntt_get_rou_from_domain(uint64_t logn, b381_fr* rou) {
  if (!rou) return B381_INVALID_POINTER;
  std::lock_guard<std::mutex> lk(g_dom_mu);
  Domain& g_dom = cur_domain();
  if (!g_dom.ready || logn > g_dom.K) return B381_INVALID_ARGUMENT;
  fr_t v;
  if (logn == 0) {
    v = fr_t{FR_ONE_INIT};
  } else if (logn == 1) {
    // omega_2 = -1 = r - R in Montgomery form
    const uint64_t m[4] = FR_MODULUS_INIT, o[4] = FR_ONE_INIT;
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)m[i] - o[i] - br;
      v.l[i] = (uint64_t)d;
      br = (uint64_t)(d >> 64) & 1;
    }
  } else {
    // omega_{2^logn} = T_{logn-1}[1]
    cudaError_t e = cudaMemcpy(&v, g_dom.table + ((1ull << (logn - 1)) - 1) + 1, sizeof(fr_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return map_cuda_error(e);
  }
  memcpy(rou, &v, sizeof(v));
  return B381_SUCCESS;
}
int b381_ntt_last_info(int* out, int cap) {
  int k = 0;
  for (; k < cap && k < 4; k++) out[k] = g_ntt_last_info[k];
  return k;
}
b381_ntt_config b381_default_ntt_config(void) {
  b381_ntt_config c;
  memset(&c, 0, sizeof(c));
  const uint64_t o[4] = FR_ONE_INIT;
  memcpy(&c.coset_gen, o, sizeof(o));
  c.batch_size = 1;
  c.ordering = B381_kNN;
  return c;
}
// reference-named flat API
int bls12_381_ntt_cuda(const b381_fr* in, int size, int dir, const b381_ntt_config* cfg, b381_fr* out) { return b381_ntt(in, size, dir, cfg, out); }
int bls12_381_field_ntt_cuda(const b381_fr* in, int size, int dir, const b381_ntt_config* cfg, b381_fr* out) { return b381_ntt(in, size, dir, cfg, out); }
int bls12_381_ntt_init_domain_cuda(const b381_fr* r, const b381_ntt_init_domain_config* c) { return b381_ntt_init_domain(r, c); }
int bls12_381_field_ntt_init_domain_cuda(const b381_fr* r, const b381_ntt_init_domain_config* c) { return b381_ntt_init_domain(r, c); }
int bls12_381_ntt_release_domain_cuda(void) { return b381_ntt_release_domain(); }
int bls12_381_field_ntt_release_domain_cuda(void) { return b381_ntt_release_domain(); }
int bls12_381_coset_ntt_cuda(const b381_fr* in, int size, int dir, const b381_fr* coset_gen, const b381_ntt_config* cfg, b381_fr* out) {
  if (!cfg || !coset_gen) return B381_INVALID_POINTER;
  b381_ntt_config c = *cfg;
  c.coset_gen = *coset_gen;
  return b381_ntt(in, size, dir, &c, out);
}
}
