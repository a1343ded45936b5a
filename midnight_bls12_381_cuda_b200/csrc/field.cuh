// BLS12-381 field layer for sm_100a: thin value-semantic wrappers over the generated
// inline-PTX routines in field_ptx.cuh, plus Fq2 = Fq[u]/(u^2+1).
//
// Replaces (behaviourally) the reference's Field<Config> free functions
//   field_add/sub/neg/mul/sqr/inv/to_montgomery/from_montgomery
//   bls12-381/include/field.cuh:389-928  and  fq2_* bls12-381/include/point.cuh:131-225.
// Contract kept from the reference: little-endian u64 limbs, Montgomery domain with
// R = 2^384 (Fq) / 2^256 (Fr), every result canonical (< modulus), inv(0) = 0.
#pragma once
#ifdef B381_HOST_TEST
// CPU-only unit-test build of the per-thread device code (tests/host/); never shipped.
#include "field_host_shim.h"
#include "field_consts.h"
#define B381_DI inline
#define B381_HD inline
#else
#include "field_ptx.cuh"
#include "raw_ptx.cuh"
#define B381_DI __device__ __forceinline__
#define B381_HD __host__ __device__ __forceinline__
#endif

#include "inv_bingcd.cuh"

namespace b381 {

// ------------------------------------------------------------------ constants
B381_DI fq_t fq_modulus() { return fq_t{FQ_MODULUS_INIT}; }
B381_DI fr_t fr_modulus() { return fr_t{FR_MODULUS_INIT}; }

template <class F> B381_DI F zero();
template <class F> B381_DI F one();
template <> B381_DI fq_t zero<fq_t>() { return fq_t{{0, 0, 0, 0, 0, 0}}; }
template <> B381_DI fr_t zero<fr_t>() { return fr_t{{0, 0, 0, 0}}; }
template <> B381_DI fq_t one<fq_t>() { return fq_t{FQ_ONE_INIT}; }
template <> B381_DI fr_t one<fr_t>() { return fr_t{FR_ONE_INIT}; }

// ------------------------------------------------------------------ Fq / Fr
B381_DI fq_t mul(const fq_t& a, const fq_t& b) { fq_t r; fq_mul_raw(r, a, b); return r; }
B381_DI fq_t sqr(const fq_t& a) { fq_t r; fq_sqr_raw(r, a); return r; }
B381_DI fq_t add(const fq_t& a, const fq_t& b) { fq_t r; fq_add_raw(r, a, b); return r; }
B381_DI fq_t sub(const fq_t& a, const fq_t& b) { fq_t r; fq_sub_raw(r, a, b); return r; }
B381_DI fq_t neg(const fq_t& a) { fq_t r; fq_neg_raw(r, a); return r; }
B381_DI fq_t dbl(const fq_t& a) { fq_t r; fq_dbl_raw(r, a); return r; }

B381_DI fr_t mul(const fr_t& a, const fr_t& b) { fr_t r; fr_mul_raw(r, a, b); return r; }
B381_DI fr_t sqr(const fr_t& a) { fr_t r; fr_sqr_raw(r, a); return r; }
B381_DI fr_t add(const fr_t& a, const fr_t& b) { fr_t r; fr_add_raw(r, a, b); return r; }
B381_DI fr_t sub(const fr_t& a, const fr_t& b) { fr_t r; fr_sub_raw(r, a, b); return r; }
B381_DI fr_t neg(const fr_t& a) { fr_t r; fr_neg_raw(r, a); return r; }
B381_DI fr_t dbl(const fr_t& a) { fr_t r; fr_dbl_raw(r, a); return r; }

B381_DI bool is_zero(const fq_t& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5]) == 0; }
B381_DI bool is_zero(const fr_t& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
B381_DI bool eq(const fq_t& a, const fq_t& b) {
  return ((a.l[0] ^ b.l[0]) | (a.l[1] ^ b.l[1]) | (a.l[2] ^ b.l[2]) | (a.l[3] ^ b.l[3]) |
          (a.l[4] ^ b.l[4]) | (a.l[5] ^ b.l[5])) == 0;
}
B381_DI bool eq(const fr_t& a, const fr_t& b) {
  return ((a.l[0] ^ b.l[0]) | (a.l[1] ^ b.l[1]) | (a.l[2] ^ b.l[2]) | (a.l[3] ^ b.l[3])) == 0;
}

// Montgomery <-> standard (reference: field.cuh:906-928)
B381_DI fq_t to_mont(const fq_t& a) { return mul(a, fq_t{FQ_R2_INIT}); }
B381_DI fr_t to_mont(const fr_t& a) { return mul(a, fr_t{FR_R2_INIT}); }
B381_DI fq_t from_mont(const fq_t& a) { fq_t o = {{1, 0, 0, 0, 0, 0}}; return mul(a, o); }
B381_DI fr_t from_mont(const fr_t& a) { fr_t o = {{1, 0, 0, 0}}; return mul(a, o); }

// a^(m-2) by plain square-and-multiply; only ever runs on single-thread tails
// (one inversion per MSM, n^-1 per NTT domain).  inv(0) = 0 as in field.cuh:750-900.
template <class F, int N64>
B381_DI F pow_m_minus_2(const F& a, const uint64_t (&m)[N64]) {
  uint64_t e[N64];
#pragma unroll
  for (int i = 0; i < N64; i++) e[i] = m[i];
  e[0] -= 2;  // both moduli end in ...01 / ...ab, no borrow
  F r = one<F>();
#pragma unroll 1
  for (int i = N64 * 64 - 1; i >= 0; i--) {
    r = sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, a);
  }
  return r;
}
B381_DI fq_t inv(const fq_t& a) { const uint64_t m[6] = FQ_MODULUS_INIT; return pow_m_minus_2<fq_t, 6>(a, m); }
B381_DI fr_t inv(const fr_t& a) { const uint64_t m[4] = FR_MODULUS_INIT; return pow_m_minus_2<fr_t, 4>(a, m); }

// ---- variable-time inversion (Kaliski almost-inverse + 3 Montgomery products) -------------------
// Used by the batched inversions of the affine bucket pre-reduction (msm_batch.cuh): ~560 rounds of
// 384-bit shift/subtract instead of the ~760 dependent Montgomery products of a^(m-2).  Public data
// only (MSM bases).  Phase 1 leaves x = A^-1 * 2^k (mod p), 381 <= k <= 762, for the integer A = a*R
// held in `a`; a^-1 * R = x * 2^(768-k), applied as mont(mont(mont(x, R^3), 2^e2), 2^e3), e2+e3 = 768-k.
//
// The round is written for SIMT: u stays odd, only v is halved, and the "v < u" case swaps the roles
// of (u,r) and (v,s) with masks, so every lane of a warp runs the same instructions whatever its
// value -- the textbook four-way branch made a warp of 32 different inversions execute all four arms
// every round (ncu: 211 k instructions per warp-inversion).  Invariants, sigma = +-1 flipping on a swap:
//   u*s + v*r = p,   a*r = -sigma*u*2^k,   a*s = sigma*v*2^k  (mod p).
B381_HD uint64_t raw6_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
#if defined(__CUDA_ARCH__)
  return raw6_add_ptx(r, a, b);
#endif
  unsigned __int128 c = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) { c += (unsigned __int128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
B381_HD uint64_t raw6_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
#if defined(__CUDA_ARCH__)
  return raw6_sub_ptx(r, a, b);
#endif
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    unsigned __int128 d = (unsigned __int128)a[i] - b[i] - br;
    r[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}
// m ? a : b for an all-ones / all-zeros mask: two LOP3 on the device (the plain expression becomes SEL + LOP3 pairs)
B381_HD uint64_t sel64(uint64_t m, uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t rl, rh;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(rl) : "r"((uint32_t)a), "r"((uint32_t)b), "r"((uint32_t)m));
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(rh) : "r"((uint32_t)(a >> 32)), "r"((uint32_t)(b >> 32)), "r"((uint32_t)(m >> 32)));
  return ((uint64_t)rh << 32) | rl;
#else
  return (a & m) | (b & ~m);
#endif
}
B381_HD void raw6_shr1(uint64_t* a) {
#pragma unroll
  for (int i = 0; i < 5; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
  a[5] >>= 1;
}
B381_HD void raw6_shl1(uint64_t* a) {
#pragma unroll
  for (int i = 5; i > 0; i--) a[i] = (a[i] << 1) | (a[i - 1] >> 63);
  a[0] <<= 1;
}
// Variable-time inversion of a Montgomery-form element: the integer held in `a` is A = a*R; the binary GCD of
// inv_bingcd.cuh gives A^-1 (mod p) and one Montgomery product with R^3 turns it into a^-1 * R.  inv(0) = 0.
// (Round 1's Kaliski almost-inverse below is kept for A/B runs: -DB381_INV_KALISKI.)
#ifndef B381_INV_KALISKI
B381_DI fq_t inv_vartime(const fq_t& a) {
  const uint64_t P[6] = FQ_MODULUS_INIT;
  uint32_t y[12], m[12], o[12];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    y[2 * i] = (uint32_t)a.l[i]; y[2 * i + 1] = (uint32_t)(a.l[i] >> 32);
    m[2 * i] = (uint32_t)P[i];   m[2 * i + 1] = (uint32_t)(P[i] >> 32);
  }
  bingcd_inverse<12>(y, m, FQ_INV32, kFqInvRounds, o);
  fq_t x;
#pragma unroll
  for (int i = 0; i < 6; i++) x.l[i] = ((uint64_t)o[2 * i + 1] << 32) | o[2 * i];
  return mul(x, fq_t{FQ_R3_INIT});
}
#else
B381_DI fq_t inv_vartime(const fq_t& a) {
  if (is_zero(a)) return a;
  const uint64_t P[6] = FQ_MODULUS_INIT;
  uint64_t u[6], v[6], r[6], s[6], t1[6], t2[6], ss[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { u[i] = P[i]; v[i] = a.l[i]; r[i] = 0; s[i] = 0; }
  s[0] = 1;
  uint32_t k = 0;
  uint64_t sigma_neg = 0;
#pragma unroll 1
  while ((v[0] | v[1] | v[2] | v[3] | v[4] | v[5]) != 0) {
    const uint64_t odd = 0 - (v[0] & 1);
    const uint64_t lt = raw6_sub(t1, v, u);      // t1 = v - u, borrow <=> v < u
    raw6_sub(t2, u, v);                          // t2 = u - v
    raw6_add(ss, r, s);
    const uint64_t swp = odd & (0 - lt);
#pragma unroll
    for (int i = 0; i < 6; i++) {
      uint64_t vn = sel64(swp, t2[i], t1[i]);
      uint64_t vi = sel64(odd, vn, v[i]);
      u[i] = sel64(swp, v[i], u[i]);
      uint64_t rn = sel64(swp, s[i], r[i]);
      s[i] = sel64(odd, ss[i], s[i]);
      r[i] = rn;
      v[i] = vi;
    }
    raw6_shr1(v);
    raw6_shl1(r);
    sigma_neg ^= swp & 1;
    k++;
  }
  if (raw6_sub(t1, r, P) == 0) {
#pragma unroll
    for (int i = 0; i < 6; i++) r[i] = t1[i];
  }
  fq_t x;
  if (sigma_neg) {
#pragma unroll
    for (int i = 0; i < 6; i++) x.l[i] = r[i];
  } else {
    raw6_sub(x.l, P, r);
  }
  uint32_t e = 768u - k, e2 = e > 380u ? 380u : e, e3 = e - e2;
  fq_t c2 = zero<fq_t>(), c3 = zero<fq_t>();
#pragma unroll
  for (int i = 0; i < 6; i++) {
    if ((e2 >> 6) == (uint32_t)i) c2.l[i] = 1ull << (e2 & 63);
    if ((e3 >> 6) == (uint32_t)i) c3.l[i] = 1ull << (e3 & 63);
  }
  x = mul(x, fq_t{FQ_R3_INIT});
  x = mul(x, c2);
  return mul(x, c3);
}
#endif

// r = a^e for a 64-bit exponent (twiddle / coset power setup)
template <class F>
B381_DI F pow_u64(const F& a, uint64_t e) {
  F r = one<F>();
  F b = a;
#pragma unroll 1
  while (e) {
    if (e & 1) r = mul(r, b);
    b = sqr(b);
    e >>= 1;
  }
  return r;
}

// ------------------------------------------------------------------ Fq2
struct fq2_t { fq_t c0, c1; };
template <> B381_DI fq2_t zero<fq2_t>() { return fq2_t{zero<fq_t>(), zero<fq_t>()}; }
template <> B381_DI fq2_t one<fq2_t>() { return fq2_t{one<fq_t>(), zero<fq_t>()}; }

B381_DI fq2_t add(const fq2_t& a, const fq2_t& b) { return fq2_t{add(a.c0, b.c0), add(a.c1, b.c1)}; }
B381_DI fq2_t sub(const fq2_t& a, const fq2_t& b) { return fq2_t{sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
B381_DI fq2_t neg(const fq2_t& a) { return fq2_t{neg(a.c0), neg(a.c1)}; }
B381_DI fq2_t dbl(const fq2_t& a) { return fq2_t{dbl(a.c0), dbl(a.c1)}; }
B381_DI fq2_t mul(const fq2_t& a, const fq2_t& b) {
  // Karatsuba, 3 Fq products
  fq_t v0 = mul(a.c0, b.c0);
  fq_t v1 = mul(a.c1, b.c1);
  fq_t s = mul(add(a.c0, a.c1), add(b.c0, b.c1));
  return fq2_t{sub(v0, v1), sub(sub(s, v0), v1)};
}
B381_DI fq2_t sqr(const fq2_t& a) {
  // (a0+a1)(a0-a1) + 2 a0 a1 u, 2 Fq products
  fq_t t = mul(a.c0, a.c1);
  fq_t c0 = mul(add(a.c0, a.c1), sub(a.c0, a.c1));
  return fq2_t{c0, dbl(t)};
}
B381_DI bool is_zero(const fq2_t& a) { return is_zero(a.c0) && is_zero(a.c1); }
B381_DI bool eq(const fq2_t& a, const fq2_t& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
B381_DI fq2_t inv(const fq2_t& a) {
  fq_t n = inv(add(sqr(a.c0), sqr(a.c1)));
  return fq2_t{mul(a.c0, n), neg(mul(a.c1, n))};
}
B381_DI fq2_t inv_vartime(const fq2_t& a) {
  fq_t n = inv_vartime(add(sqr(a.c0), sqr(a.c1)));
  return fq2_t{mul(a.c0, n), neg(mul(a.c1, n))};
}
B381_DI fq2_t to_mont(const fq2_t& a) { return fq2_t{to_mont(a.c0), to_mont(a.c1)}; }
B381_DI fq2_t from_mont(const fq2_t& a) { return fq2_t{from_mont(a.c0), from_mont(a.c1)}; }

}  // namespace b381
