// CPU-only driver for the per-thread bodies of csrc/glv.cuh (TEST INFRASTRUCTURE).  Usage:
//   glv_host_sim decompose <infile>         infile = n x 32 B canonical scalars; prints "k1 k2" (hex, 128 bit each) per line
//   glv_host_sim mul <glv 0|1> <infile>     infile = n x (96 B Montgomery affine G1 point + 32 B canonical scalar);
//                                           prints the 96-byte Montgomery affine result (hex) per line
//   glv_host_sim sub <g1|g2> <infile>       infile = n x (96|192) B Montgomery affine points; prints 0/1 per line
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define B381_HOST_TEST 1
#include "glv.cuh"
using namespace b381;

static std::vector<unsigned char> slurp(const char* path) {
  std::vector<unsigned char> v;
  FILE* f = fopen(path, "rb");
  if (!f) exit(3);
  unsigned char buf[4096];
  size_t k;
  while ((k = fread(buf, 1, sizeof(buf), f)) > 0) v.insert(v.end(), buf, buf + k);
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  if (!strcmp(argv[1], "decompose")) {
    auto d = slurp(argv[2]);
    for (size_t i = 0; i + 32 <= d.size(); i += 32) {
      fr_t k;
      memcpy(&k, &d[i], 32);
      uint64_t k1[2], k2[2];
      glv_decompose(k, k1, k2);
      printf("%016llx%016llx %016llx%016llx\n", (unsigned long long)k1[1], (unsigned long long)k1[0], (unsigned long long)k2[1],
             (unsigned long long)k2[0]);
    }
    return 0;
  }
  if (!strcmp(argv[1], "mul") && argc >= 4) {
    const bool glv = atoi(argv[2]) != 0;
    auto d = slurp(argv[3]);
    for (size_t i = 0; i + 128 <= d.size(); i += 128) {
      g1_affine p;
      fr_t k;
      memcpy(&p, &d[i], 96);
      memcpy(&k, &d[i + 96], 32);
      g1_affine a = xyzz_to_affine(glv ? g1_mul_glv(p, k) : g1_mul_window(p, k));
      const unsigned char* q = (const unsigned char*)&a;
      for (size_t j = 0; j < 96; j++) printf("%02x", q[j]);
      printf("\n");
    }
    return 0;
  }
  if (!strcmp(argv[1], "sub") && argc >= 4) {
    const bool g2 = argv[2][1] == '2';
    auto d = slurp(argv[3]);
    const size_t sz = g2 ? 192 : 96;
    for (size_t i = 0; i + sz <= d.size(); i += sz) {
      if (g2) { g2_affine p; memcpy(&p, &d[i], sz); printf("%d\n", g2_in_subgroup(p) ? 1 : 0); }
      else { g1_affine p; memcpy(&p, &d[i], sz); printf("%d\n", g1_in_subgroup(p) ? 1 : 0); }
    }
    return 0;
  }
  return 1;
}
