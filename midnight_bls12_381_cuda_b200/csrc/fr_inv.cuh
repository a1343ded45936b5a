// Variable-time inversion in Fr: the 4-limb twin of field.cuh's inv_vartime (same SIMT-uniform Kaliski
// almost-inverse; see the derivation there).  Phase 1 leaves x = A^-1 * 2^k (mod r), 255 <= k <= 510, for the
// integer A = a*R held in `a`; a^-1 * R = x * 2^(512-k) = mont(mont(mont(x, R^3), 2^e2), 2^e3), e2 + e3 = 512 - k,
// both single-bit constants below r (e2 <= 254).  inv(0) = 0 like the reference's field_inv
// (bls12-381/include/field.cuh:750-900).  Used by the batched Fr inversion in vecops.cu.
#pragma once
#include "field.cuh"

namespace b381 {

template <int N>
B381_HD uint64_t rawn_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
#if defined(__CUDA_ARCH__)
  if (N == 4) return raw4_add_ptx(r, a, b);
#endif
  unsigned __int128 c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { c += (unsigned __int128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
template <int N>
B381_HD uint64_t rawn_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
#if defined(__CUDA_ARCH__)
  if (N == 4) return raw4_sub_ptx(r, a, b);
#endif
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    unsigned __int128 d = (unsigned __int128)a[i] - b[i] - br;
    r[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}

#ifndef B381_INV_KALISKI
// binary GCD on the integer A = a*R (inv_bingcd.cuh), then one Montgomery product with R^3: a^-1 * R; inv(0) = 0
B381_DI fr_t inv_vartime(const fr_t& a) {
  const uint64_t P[4] = FR_MODULUS_INIT;
  uint32_t y[8], m[8], o[8];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    y[2 * i] = (uint32_t)a.l[i]; y[2 * i + 1] = (uint32_t)(a.l[i] >> 32);
    m[2 * i] = (uint32_t)P[i];   m[2 * i + 1] = (uint32_t)(P[i] >> 32);
  }
  bingcd_inverse<8>(y, m, FR_INV32, kFrInvRounds, o);
  fr_t x;
#pragma unroll
  for (int i = 0; i < 4; i++) x.l[i] = ((uint64_t)o[2 * i + 1] << 32) | o[2 * i];
  return mul(x, fr_t{FR_R3_INIT});
}
#else
B381_DI fr_t inv_vartime(const fr_t& a) {
  if (is_zero(a)) return a;
  constexpr int N = 4;
  const uint64_t P[N] = FR_MODULUS_INIT;
  uint64_t u[N], v[N], r[N], s[N], t1[N], t2[N], ss[N];
#pragma unroll
  for (int i = 0; i < N; i++) { u[i] = P[i]; v[i] = a.l[i]; r[i] = 0; s[i] = 0; }
  s[0] = 1;
  uint32_t k = 0;
  uint64_t sigma_neg = 0;
#pragma unroll 1
  while ((v[0] | v[1] | v[2] | v[3]) != 0) {
    const uint64_t odd = 0 - (v[0] & 1);
    const uint64_t lt = rawn_sub<N>(t1, v, u);     // t1 = v - u, borrow <=> v < u
    rawn_sub<N>(t2, u, v);
    rawn_add<N>(ss, r, s);
    const uint64_t swp = odd & (0 - lt);
#pragma unroll
    for (int i = 0; i < N; i++) {
      uint64_t vn = sel64(swp, t2[i], t1[i]);
      uint64_t vi = sel64(odd, vn, v[i]);
      u[i] = sel64(swp, v[i], u[i]);
      uint64_t rn = sel64(swp, s[i], r[i]);
      s[i] = sel64(odd, ss[i], s[i]);
      r[i] = rn;
      v[i] = vi;
    }
#pragma unroll
    for (int i = 0; i < N - 1; i++) v[i] = (v[i] >> 1) | (v[i + 1] << 63);
    v[N - 1] >>= 1;
#pragma unroll
    for (int i = N - 1; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
    r[0] <<= 1;
    sigma_neg ^= swp & 1;
    k++;
  }
  if (rawn_sub<N>(t1, r, P) == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = t1[i];
  }
  fr_t x;
  if (sigma_neg) {
#pragma unroll
    for (int i = 0; i < N; i++) x.l[i] = r[i];
  } else {
    rawn_sub<N>(x.l, P, r);
  }
  uint32_t e = 512u - k, e2 = e > 254u ? 254u : e, e3 = e - e2;
  fr_t c2 = zero<fr_t>(), c3 = zero<fr_t>();
#pragma unroll
  for (int i = 0; i < N; i++) {
    if ((e2 >> 6) == (uint32_t)i) c2.l[i] = 1ull << (e2 & 63);
    if ((e3 >> 6) == (uint32_t)i) c3.l[i] = 1ull << (e3 & 63);
  }
  x = mul(x, fr_t{FR_R3_INIT});
  x = mul(x, c2);
  return mul(x, c3);
}
#endif

}  // namespace b381
