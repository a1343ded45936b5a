// Host-side plumbing shared by the .cu translation units: error mapping, stream-ordered
// scratch memory and residency staging.  (The reference cudaMalloc/cudaFree's up to 11 buffers
// per MSM call and maps every failure to ALLOCATION_FAILED -- src/curve/msm_kernels.cu:705-719,
// include/icicle_types.cuh:41-45; here scratch comes from the stream-ordered pool, so a call
// never implicitly synchronises the device.)
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/b381.h"

namespace b381 {

inline int map_cuda_error(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return B381_SUCCESS;
    case cudaErrorMemoryAllocation: return B381_OUT_OF_MEMORY;
    case cudaErrorInvalidValue: return B381_INVALID_ARGUMENT;
    case cudaErrorInvalidDevice:
    case cudaErrorNoDevice:
    case cudaErrorInsufficientDriver: return B381_INVALID_DEVICE;
    default: return B381_UNKNOWN_ERROR;
  }
}

#define B381_CUDA_TRY(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      if (getenv("B381_DEBUG"))                                                        \
        fprintf(stderr, "[b381] %s:%d %s -> %s\n", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return _e;                                                                       \
    }                                                                                  \
  } while (0)

// Keep freed scratch cached in the default pool instead of returning it to the driver.
inline void configure_pool_once() {
  static bool done[64] = {};                 // per device: a single process may drive several GPUs
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    // Never satisfy an allocation on stream B out of memory whose free is still pending on stream A by making B WAIT
    // for A: two calls in flight on two streams (the async handles of the host API) would run one after the other.
    // With 180 GB the pool simply grows to the working set of both.
    int off = 0;
    cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
  }
  done[dev] = true;
}

// RAII bag of stream-ordered allocations; everything is released (asynchronously, in stream
// order) when the bag dies, so is_async callers return without a host sync.
class Scratch {
 public:
  explicit Scratch(cudaStream_t s) : s_(s) { configure_pool_once(); }
  ~Scratch() {
    for (void* p : ptrs_) cudaFreeAsync(p, s_);
  }
  template <class T>
  cudaError_t alloc(T** out, size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(&p, bytes, s_);
    if (e != cudaSuccess) return e;
    ptrs_.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return cudaSuccess;
  }
  cudaStream_t stream() const { return s_; }

 private:
  cudaStream_t s_;
  std::vector<void*> ptrs_;
};

// Device view of a caller buffer: used in place when it already lives on the device, else
// copied into scratch on the stream.
template <class T>
cudaError_t stage_in(Scratch& sc, const T* src, size_t count, bool on_device, const T** out) {
  if (on_device) {
    *out = src;
    return cudaSuccess;
  }
  T* d = nullptr;
  cudaError_t e = sc.alloc(&d, count);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, sc.stream());
  *out = d;
  return e;
}

// NVTX ranges (the reference only has Rust `tracing` spans around its calls, core/msm.rs:538-574,
// core/ntt.rs:509-541).  NVTX3 is header-only and resolves the tool's injection library at run time, so there is no link
// dependency and a range costs a few nanoseconds when no profiler is attached.  Domain "b381"; ranges: one per C-ABI
// call (with the size in the name) and one per MSM phase, so a timeline / ncu --nvtx-include filter can address them.
struct TraceRange {
  static nvtxDomainHandle_t domain() {
    static nvtxDomainHandle_t d = nvtxDomainCreateA("b381");
    return d;
  }
  explicit TraceRange(const char* name, long long arg = -1) {
    char buf[96];
    if (arg >= 0) { snprintf(buf, sizeof(buf), "%s n=%lld", name, arg); name = buf; }
    nvtxEventAttributes_t a = {};
    a.version = NVTX_VERSION;
    a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
    a.messageType = NVTX_MESSAGE_TYPE_ASCII;
    a.message.ascii = name;
    nvtxDomainRangePushEx(domain(), &a);
  }
  ~TraceRange() { nvtxDomainRangePop(domain()); }
  TraceRange(const TraceRange&) = delete;
  TraceRange& operator=(const TraceRange&) = delete;
};

inline unsigned grid_for(size_t threads, unsigned block) { return (unsigned)((threads + block - 1) / block); }

}  // namespace b381
