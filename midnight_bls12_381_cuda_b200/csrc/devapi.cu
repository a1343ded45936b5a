// Device plumbing behind the C ABI: pure CUDA-runtime forwarding, no kernels.
// Mirrors what CudaDeviceAPI offers ICICLE (bls12-381/src/device/cuda_device_api.cu:38-149).
#include <cstring>

#include "common.cuh"

using namespace b381;
extern "C" {
int b381_device_count(int* count) { return map_cuda_error(cudaGetDeviceCount(count)); }
int b381_set_device(int id) { return map_cuda_error(cudaSetDevice(id)); }
int b381_malloc(void** p, size_t n) { return map_cuda_error(cudaMalloc(p, n)); }
int b381_malloc_async(void** p, size_t n, void* s) { configure_pool_once(); return map_cuda_error(cudaMallocAsync(p, n, (cudaStream_t)s)); }
int b381_free(void* p) { return map_cuda_error(cudaFree(p)); }
int b381_free_async(void* p, void* s) { return map_cuda_error(cudaFreeAsync(p, (cudaStream_t)s)); }
int b381_memset(void* p, int v, size_t n) { return map_cuda_error(cudaMemset(p, v, n)); }
int b381_copy_to_device(void* d, const void* s, size_t n) { return map_cuda_error(cudaMemcpy(d, s, n, cudaMemcpyHostToDevice)); }
int b381_copy_to_host(void* d, const void* s, size_t n) { return map_cuda_error(cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)); }
int b381_copy_to_device_async(void* d, const void* s, size_t n, void* st) { return map_cuda_error(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, (cudaStream_t)st)); }
int b381_copy_to_host_async(void* d, const void* s, size_t n, void* st) { return map_cuda_error(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, (cudaStream_t)st)); }
int b381_copy_device_to_device(void* d, const void* s, size_t n, void* st) { return map_cuda_error(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, (cudaStream_t)st)); }
int b381_host_alloc_pinned(void** p, size_t n) { return map_cuda_error(cudaHostAlloc(p, n, cudaHostAllocDefault)); }
int b381_host_free_pinned(void* p) { return map_cuda_error(cudaFreeHost(p)); }
int b381_stream_create(void** s) {
  cudaStream_t st;
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e != cudaSuccess) return B381_STREAM_CREATION_FAILED;
  *s = st;
  return B381_SUCCESS;
}
int b381_stream_destroy(void* s) { return cudaStreamDestroy((cudaStream_t)s) == cudaSuccess ? B381_SUCCESS : B381_STREAM_DESTRUCTION_FAILED; }
int b381_stream_synchronize(void* s) { return cudaStreamSynchronize((cudaStream_t)s) == cudaSuccess ? B381_SUCCESS : B381_SYNCHRONIZATION_FAILED; }
int b381_device_synchronize(void) { return cudaDeviceSynchronize() == cudaSuccess ? B381_SUCCESS : B381_SYNCHRONIZATION_FAILED; }
// Peer-visible device buffers for the fused exchange of the four-step NTT (one process per GPU): a plain cudaMalloc
// allocation, its 64-byte IPC handle for the other ranks, and the mapping of a peer's handle into this process.
int b381_ipc_alloc(size_t bytes, void** ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!ptr || !handle) return B381_INVALID_POINTER;
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return map_cuda_error(e);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return map_cuda_error(e); }
  memcpy(handle, &h, 64);
  return B381_SUCCESS;
}
int b381_ipc_open(const unsigned char handle[64], void** ptr) {
  if (!ptr || !handle) return B381_INVALID_POINTER;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  return map_cuda_error(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
}
int b381_ipc_close(void* ptr) { return map_cuda_error(cudaIpcCloseMemHandle(ptr)); }
const char* b381_version(void) { return "b381-cuda-b200 0.1 (sm_100a)"; }
}
