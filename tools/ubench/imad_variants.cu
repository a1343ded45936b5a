// Micro-benchmarks: issue cost of the IMAD.WIDE flavours that a Montgomery multiplier can be built from.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4000
// V0: plain mad.wide (no carry)          V1: chained carry (mad.lo.cc/madc.hi.cc ... -> IMAD.WIDE.U32.X)
// V2: carry-out only pairs + addc sink     V3: mad.lo + mad.hi separate (IMAD + IMAD.HI)
template <int V> __global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
  uint32_t a[6], b = seed * 3 + blockIdx.x;
  uint32_t lo[6], hi[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { a[i] = seed + threadIdx.x * (i + 1); lo[i] = i; hi[i] = i + threadIdx.x; }
  uint32_t sink = 0;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    if (V == 0) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 6; i++) {
          uint64_t acc = ((uint64_t)hi[i] << 32) | lo[i];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[i]), "r"(b + r));
          lo[i] = (uint32_t)acc; hi[i] = (uint32_t)(acc >> 32);
        }
    } else if (V == 1) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        asm volatile(
            "mad.lo.cc.u32 %0, %12, %18, %0;\n\tmadc.hi.cc.u32 %1, %12, %18, %1;\n\t"
            "madc.lo.cc.u32 %2, %13, %18, %2;\n\tmadc.hi.cc.u32 %3, %13, %18, %3;\n\t"
            "madc.lo.cc.u32 %4, %14, %18, %4;\n\tmadc.hi.cc.u32 %5, %14, %18, %5;\n\t"
            "madc.lo.cc.u32 %6, %15, %18, %6;\n\tmadc.hi.cc.u32 %7, %15, %18, %7;\n\t"
            "madc.lo.cc.u32 %8, %16, %18, %8;\n\tmadc.hi.cc.u32 %9, %16, %18, %9;\n\t"
            "madc.lo.cc.u32 %10, %17, %18, %10;\n\tmadc.hi.cc.u32 %11, %17, %18, %11;\n\t"
            : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
              "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(b + r));
      }
    } else if (V == 2) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 6; i++)
          asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                       : "+r"(lo[i]), "+r"(hi[i]), "+r"(sink) : "r"(a[i]), "r"(b + r));
    } else {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 6; i++) {
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a[i]), "r"(b + r));
          asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(a[i]), "r"(b + r));
        }
    }
  }
  uint32_t s = sink;
#pragma unroll
  for (int i = 0; i < 6; i++) s ^= lo[i] ^ hi[i];
  if (s == 0x12345679u) out[0] = s;
}
template <int V> void run(const char* name) {
  uint32_t* d; cudaMalloc(&d, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = 148 * 8;
  k<V><<<blocks, 256>>>(d, 7); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<V><<<blocks, 256>>>(d, 7); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double prods = (double)blocks * 256 * 24.0 * ITERS;
  printf("%-28s %.3f ms  %.3e wide-products/s  (%.2f per clk per SM @1.965GHz)\n", name, ms, prods / (ms * 1e-3), prods / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  run<0>("mad.wide (no carry)");
  run<1>("carry chain (.X in+out)");
  run<2>("carry-out only + addc");
  run<3>("mad.lo + mad.hi separate");
  return 0;
}
