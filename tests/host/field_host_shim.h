// TEST-ONLY host stand-in for the generated inline-PTX routines (field_ptx.cuh).
// Compiled only into tests/host/* harnesses (g++ -DB381_HOST_TEST) so that the device headers'
// curve formulas and per-thread MSM/NTT bodies can be exercised on a CPU-only box.
// It is never part of the shipped libraries: csrc/field.cuh includes it only under
// B381_HOST_TEST, and build.py never defines that macro.
#pragma once
#include <cstdint>
#include <cstring>

struct fq_t { uint64_t l[6]; };
struct fr_t { uint64_t l[4]; };

namespace host_shim {
typedef unsigned __int128 u128;

template <int N> struct mod_t { uint64_t m[N]; uint64_t inv; };
static const mod_t<6> FQM = {{0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                              0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
                             0x89f3fffcfffcfffdull};
static const mod_t<4> FRM = {{0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                              0x73eda753299d7d48ull},
                             0xfffffffeffffffffull};

template <int N> inline bool geq(const uint64_t* a, const uint64_t* b) {
  for (int i = N - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
  return true;
}
template <int N> inline uint64_t add_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 c = 0;
  for (int i = 0; i < N; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
template <int N> inline uint64_t sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t br = 0;
  for (int i = 0; i < N; i++) {
    u128 d = (u128)a[i] - b[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}
template <int N> inline void mont_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, const mod_t<N>& M) {
  uint64_t t[N + 2];
  memset(t, 0, sizeof(t));
  for (int i = 0; i < N; i++) {
    u128 c = 0;
    for (int j = 0; j < N; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * M.inv;
    c = (u128)m * M.m[0] + t[0]; c >>= 64;
    for (int j = 1; j < N; j++) { c += (u128)m * M.m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[N]; t[N - 1] = (uint64_t)c; t[N] = t[N + 1] + (uint64_t)(c >> 64);
  }
  if (t[N] || geq<N>(t, M.m)) sub_n<N>(t, t, M.m);
  memcpy(r, t, N * 8);
}
template <int N> inline void mod_add(uint64_t* r, const uint64_t* a, const uint64_t* b, const mod_t<N>& M) {
  uint64_t t[N]; uint64_t c = add_n<N>(t, a, b);
  if (c || geq<N>(t, M.m)) sub_n<N>(t, t, M.m);
  memcpy(r, t, N * 8);
}
template <int N> inline void mod_sub(uint64_t* r, const uint64_t* a, const uint64_t* b, const mod_t<N>& M) {
  uint64_t t[N]; if (sub_n<N>(t, a, b)) add_n<N>(t, t, M.m);
  memcpy(r, t, N * 8);
}
}  // namespace host_shim

#define SHIM2(F, N, MOD)                                                                             \
  inline void F##_mul_raw(F##_t& r, const F##_t& a, const F##_t& b) { host_shim::mont_mul<N>(r.l, a.l, b.l, host_shim::MOD); } \
  inline void F##_sqr_raw(F##_t& r, const F##_t& a) { host_shim::mont_mul<N>(r.l, a.l, a.l, host_shim::MOD); }               \
  inline void F##_add_raw(F##_t& r, const F##_t& a, const F##_t& b) { host_shim::mod_add<N>(r.l, a.l, b.l, host_shim::MOD); } \
  inline void F##_sub_raw(F##_t& r, const F##_t& a, const F##_t& b) { host_shim::mod_sub<N>(r.l, a.l, b.l, host_shim::MOD); } \
  inline void F##_dbl_raw(F##_t& r, const F##_t& a) { host_shim::mod_add<N>(r.l, a.l, a.l, host_shim::MOD); }               \
  inline void F##_neg_raw(F##_t& r, const F##_t& a) { F##_t z; memset(&z, 0, sizeof(z)); host_shim::mod_sub<N>(r.l, z.l, a.l, host_shim::MOD); }
SHIM2(fq, 6, FQM)
SHIM2(fr, 4, FRM)
#undef SHIM2

// CUDA vector types used by the per-thread bodies
struct uint2 { unsigned int x, y; };
struct uint4 { unsigned int x, y, z, w; };
