// L2 atomic / scattered-store throughput on random addresses (counting-sort building blocks of the MSM front end):
//   RED  : atomicAdd without return over R counters            (histogram pass)
//   ATOM : atomicAdd with return + dependent 4-byte store       (scatter pass)
// usage: atomic_probe  -- prints ops/s for R = 2^15 .. 2^24 counters, 2^28 ops, 16 independent ops per thread
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__global__ void k_probe(uint32_t* ctr, uint32_t rmask, uint32_t* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t pos[16];
#pragma unroll
  for (int w = 0; w < 16; w++) {
    uint32_t k = mix(i * 16u + w) & rmask;
    if (MODE == 0) atomicAdd(ctr + (size_t)w * (rmask + 1) + k, 1u);
    else pos[w] = atomicAdd(ctr + (size_t)w * (rmask + 1) + k, 1u);
  }
  if (MODE == 1) {
#pragma unroll
    for (int w = 0; w < 16; w++) {
      uint32_t k = mix(i * 16u + w) & rmask;
      // bucket-local cursor: position inside a slab of 2^28 / (16 R) entries per counter (wraps)
      uint32_t slab = (1u << 28) / (16u * (rmask + 1));
      out[((size_t)w * (rmask + 1) + k) * slab + (pos[w] % slab)] = i;
    }
  }
}

int main() {
  const uint32_t n = 1u << 24;
  uint32_t *ctr, *out;
  cudaMalloc(&ctr, 16ull * (1u << 24) * 4);
  cudaMalloc(&out, (1ull << 28) * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int lr = 11; lr <= 20; lr += 1) {
    uint32_t rmask = (1u << lr) - 1;
    for (int mode = 0; mode < 2; mode++) {
      float best = 1e9;
      for (int it = 0; it < 4; it++) {
        cudaMemsetAsync(ctr, 0, 16ull * (rmask + 1) * 4);
        cudaEventRecord(a);
        if (mode == 0) k_probe<0><<<n / 256, 256>>>(ctr, rmask, out, n);
        else k_probe<1><<<n / 256, 256>>>(ctr, rmask, out, n);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it && ms < best) best = ms;
      }
      printf("counters 16 x 2^%d (%6.1f MB) %s: %.3f ms for 2^28 ops = %.3e ops/s\n", lr, 16.0 * (rmask + 1) * 4 / 1e6,
             mode ? "ATOM+store" : "RED       ", best, (double)(1u << 28) / (best * 1e-3));
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
