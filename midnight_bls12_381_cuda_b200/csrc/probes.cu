// Measurement probes used by bench.py to obtain the roofline DENOMINATORS on the box itself:
//   * peak issue rate of IMAD.WIDE.U32 (the instruction every Montgomery product is made of),
//   * sustained Fq / Fr Montgomery multiplications per second with all SMs busy.
#include "common.cuh"
#include "field.cuh"

namespace b381 {

// 8 independent accumulator chains per thread, 32 IMAD.WIDE per loop trip, no memory traffic.
__global__ void __launch_bounds__(256) k_imad_probe(uint64_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint64_t acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = k + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b + r));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= acc[k];
  if (s == 0x1234567812345678ull) out[0] = s;   // never true in practice; defeats DCE
}

template <class F>
__global__ void __launch_bounds__(128) k_mul_probe(F* out, int iters) {
  F x = one<F>(), y = one<F>();
  x.l[0] ^= threadIdx.x + 2;
  y.l[0] ^= blockIdx.x + 3;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
    x = mul(x, y);
    y = mul(y, x);
  }
  if (is_zero(x) && is_zero(y)) out[0] = x;
}

static int time_kernel(void (*launch)(cudaStream_t), float* ms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch(0);  // warm-up
  cudaEventRecord(e0, 0);
  launch(0);
  cudaEventRecord(e1, 0);
  cudaError_t e = cudaEventSynchronize(e1);
  cudaEventElapsedTime(ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return map_cuda_error(e);
}

}  // namespace b381
using namespace b381;

static int g_iters;
static void* g_buf;
static int g_blocks;
static void launch_imad(cudaStream_t s) { k_imad_probe<<<g_blocks, 256, 0, s>>>((uint64_t*)g_buf, g_iters, 12345u); }
static void launch_fq(cudaStream_t s) { k_mul_probe<fq_t><<<g_blocks, 128, 0, s>>>((fq_t*)g_buf, g_iters); }
static void launch_fr(cudaStream_t s) { k_mul_probe<fr_t><<<g_blocks, 128, 0, s>>>((fr_t*)g_buf, g_iters); }

extern "C" {
int b381_bench_imad_peak(int iters, double* mads_per_s, float* ms) {
  cudaDeviceProp p;
  int dev;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return B381_INVALID_DEVICE;
  if (cudaMalloc(&g_buf, 256) != cudaSuccess) return B381_ALLOCATION_FAILED;
  g_iters = iters;
  g_blocks = p.multiProcessorCount * 8;
  int rc = time_kernel(launch_imad, ms);
  cudaFree(g_buf);
  *mads_per_s = (double)g_blocks * 256.0 * 32.0 * iters / (*ms * 1e-3);
  return rc;
}
int b381_bench_field_mul(int field, int iters, double* muls_per_s, float* ms) {
  cudaDeviceProp p;
  int dev;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return B381_INVALID_DEVICE;
  if (cudaMalloc(&g_buf, 256) != cudaSuccess) return B381_ALLOCATION_FAILED;
  g_iters = iters;
  g_blocks = p.multiProcessorCount * 16;
  int rc = time_kernel(field == 0 ? launch_fq : launch_fr, ms);
  cudaFree(g_buf);
  *muls_per_s = (double)g_blocks * 128.0 * 2.0 * iters / (*ms * 1e-3);
  return rc;
}
}
