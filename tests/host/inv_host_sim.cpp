// CPU-only driver for csrc/inv_bingcd.cuh (TEST INFRASTRUCTURE).  Usage:
//   inv_host_sim <fq|fr> <rounds|0> <infile>     infile = k x (48|32) B little-endian integers < modulus;
//   prints y^-1 mod m (plain integers, hex, little-endian bytes) per line.  rounds = 0: the shipped round count.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define B381_HOST_TEST 1
#include "field.cuh"
#include "inv_bingcd.cuh"
using namespace b381;

template <int N>
static int run(const uint64_t* mod64, uint32_t minv32, int rounds, const char* path) {
  uint32_t m[N];
  memcpy(m, mod64, 4 * N);
  FILE* f = fopen(path, "rb");
  if (!f) return 3;
  uint32_t y[N], out[N];
  while (fread(y, 4, N, f) == (size_t)N) {
    bingcd_inverse<N>(y, m, minv32, rounds, out);
    const unsigned char* p = (const unsigned char*)out;
    for (int i = 0; i < 4 * N; i++) printf("%02x", p[i]);
    printf("\n");
  }
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) return 1;
  const int rounds = atoi(argv[2]);
  if (argv[1][1] == 'q') {
    const uint64_t P[6] = FQ_MODULUS_INIT;
    return run<12>(P, FQ_INV32, rounds ? rounds : kFqInvRounds, argv[3]);
  }
  const uint64_t P[4] = FR_MODULUS_INIT;
  return run<8>(P, FR_INV32, rounds ? rounds : kFrInvRounds, argv[3]);
}
