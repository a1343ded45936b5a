// "CUDA" DeviceAPI for ICICLE's runtime (role of bls12-381/src/device/cuda_device_api.cu:38-149):
// memory, copies, streams and device selection forwarded to the CUDA runtime.  Needed so that
// DeviceVec / IcicleStream / set_device in the Rust layer work against this backend set.
#include <cuda_runtime.h>

#include "icicle_abi.h"

namespace icicle {
namespace {

eIcicleError ok_or(cudaError_t e, eIcicleError fail) { return e == cudaSuccess ? eIcicleError::SUCCESS : fail; }

class B200DeviceAPI : public DeviceAPI {
 public:
  eIcicleError set_device(const Device& device) override { return ok_or(cudaSetDevice(device.id), eIcicleError::INVALID_DEVICE); }
  eIcicleError get_device_count(int& n) const override { return ok_or(cudaGetDeviceCount(&n), eIcicleError::INVALID_DEVICE); }
  eIcicleError allocate_memory(void** p, size_t n) const override { return ok_or(cudaMalloc(p, n), eIcicleError::ALLOCATION_FAILED); }
  eIcicleError allocate_memory_async(void** p, size_t n, icicleStreamHandle s) const override {
    return ok_or(cudaMallocAsync(p, n, static_cast<cudaStream_t>(s)), eIcicleError::ALLOCATION_FAILED);
  }
  eIcicleError free_memory(void* p) const override { return ok_or(cudaFree(p), eIcicleError::DEALLOCATION_FAILED); }
  eIcicleError free_memory_async(void* p, icicleStreamHandle s) const override {
    return ok_or(cudaFreeAsync(p, static_cast<cudaStream_t>(s)), eIcicleError::DEALLOCATION_FAILED);
  }
  eIcicleError get_available_memory(size_t& total, size_t& free) const override {
    return ok_or(cudaMemGetInfo(&free, &total), eIcicleError::UNKNOWN_ERROR);
  }
  eIcicleError memset(void* p, int v, size_t n) const override { return ok_or(cudaMemset(p, v, n), eIcicleError::UNKNOWN_ERROR); }
  eIcicleError memset_async(void* p, int v, size_t n, icicleStreamHandle s) const override {
    return ok_or(cudaMemsetAsync(p, v, n, static_cast<cudaStream_t>(s)), eIcicleError::UNKNOWN_ERROR);
  }
  static cudaMemcpyKind kind(eCopyDirection d) {
    switch (d) {
      case HostToDevice: return cudaMemcpyHostToDevice;
      case DeviceToHost: return cudaMemcpyDeviceToHost;
      case DeviceToDevice: return cudaMemcpyDeviceToDevice;
      default: return cudaMemcpyHostToHost;
    }
  }
  eIcicleError copy(void* d, const void* s, size_t n, eCopyDirection dir) const override {
    return ok_or(cudaMemcpy(d, s, n, kind(dir)), eIcicleError::COPY_FAILED);
  }
  eIcicleError copy_async(void* d, const void* s, size_t n, eCopyDirection dir, icicleStreamHandle st) const override {
    return ok_or(cudaMemcpyAsync(d, s, n, kind(dir), static_cast<cudaStream_t>(st)), eIcicleError::COPY_FAILED);
  }
  eIcicleError synchronize(icicleStreamHandle s) const override {
    return ok_or(s ? cudaStreamSynchronize(static_cast<cudaStream_t>(s)) : cudaDeviceSynchronize(),
                 eIcicleError::SYNCHRONIZATION_FAILED);
  }
  eIcicleError create_stream(icicleStreamHandle* s) const override {
    // plain cudaStreamCreate like the reference (cuda_device_api.cu:127-133): ICICLE's callers may mix such a stream
    // with default-stream work and rely on the legacy implicit ordering between the two
    cudaStream_t st;
    cudaError_t e = cudaStreamCreate(&st);
    if (e == cudaSuccess) *s = st;
    return ok_or(e, eIcicleError::STREAM_CREATION_FAILED);
  }
  eIcicleError destroy_stream(icicleStreamHandle s) const override {
    return ok_or(cudaStreamDestroy(static_cast<cudaStream_t>(s)), eIcicleError::STREAM_DESTRUCTION_FAILED);
  }
  eIcicleError get_device_properties(DeviceProperties& p) const override {
    p.using_host_memory = false;
    p.num_memory_regions = 1;             // cuda_device_api.cu:141-147
    p.supports_pinned_memory = true;
    return eIcicleError::SUCCESS;
  }
};
}  // namespace
}  // namespace icicle

B381_AT_LOAD(device) {
  if (icicle::register_deviceAPI) icicle::register_deviceAPI("CUDA", std::make_shared<icicle::B200DeviceAPI>());
}
