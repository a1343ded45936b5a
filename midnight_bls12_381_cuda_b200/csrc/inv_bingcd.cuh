// Modular inversion by the optimised binary GCD of T. Pornin ("Optimized Binary GCD for Modular Inversion", 2020),
// written for SIMT: no data-dependent branch, every lane of a warp executes the same instructions.
//
// Why: the Kaliski almost-inverse it replaces (field.cuh, round 1) runs one 384-bit add/sub/select/shift step per BIT
// of the operand (~87 k instructions per warp-inversion).  Here the bit steps run on 64-bit APPROXIMATIONS of (a, b)
// (their 34 top and 30 low bits), 30 steps at a time, producing a 2x2 matrix of 31-bit factors that is then applied
// once to the full-length (a, b) and, modulo m, to the cofactors (u, v): 27 rounds of ~900 instructions for Fq.
// The inversion sits on every latency-bound path of the MSM (one per thread of k_msm_invert_totals on every affine
// level, one per result in k_msm_encode) and in the batched Fr inversion of vecops.cu.
//
// Invariants (all values as integers; y the operand, m the odd modulus):  a = u*y (mod m),  b = v*y (mod m),
// a, b >= 0.  A round computes f0, g0, f1, g1 with |f0| + |g0| <= 2^30, |f1| + |g1| <= 2^30 such that
//   a' = (a f0 + b g0) / 2^30,   b' = (a f1 + b g1) / 2^30     (exact divisions; a sign is moved into the factors)
//   u' = (u f0 + v g0) / 2^30,   v' = (u f1 + v g1) / 2^30     (mod m: one Montgomery-style step with 2^30)
// and len(a) + len(b) shrinks by >= 30 bits per round, so ceil((2 len(m) - 1) / 30) rounds end with a = 0, b = gcd = 1,
// v = y^-1.  Further rounds leave (b, v) unchanged, so a fixed round count with one spare round is safe.  y = 0 gives 0.
//
// Replaces (behaviourally) field_inv, bls12-381/include/field.cuh:750-900 (Fermat: a^(p-2)).
#pragma once
#include <cstdint>

#ifndef B381_HD
#define B381_HD inline
#endif

namespace b381 {

constexpr int kBingcdStep = 30;      // inner steps per round = bits divided out per round
constexpr int kFqInvRounds = (2 * 381 - 1 + kBingcdStep - 1) / kBingcdStep + 1;     // 27
constexpr int kFrInvRounds = (2 * 255 - 1 + kBingcdStep - 1) / kBingcdStep + 1;     // 18

// out = (a*f + b*g) >> 30 as a signed number: low N limbs in two's complement, returns true when negative
template <int N>
B381_HD bool bingcd_lincomb(const uint32_t* a, const uint32_t* b, int32_t f, int32_t g, uint32_t* out) {
  uint32_t r[N + 1];
  int64_t acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    acc += (int64_t)(uint64_t)a[i] * f + (int64_t)(uint64_t)b[i] * g;      // |a f + b g| <= 2^32 * 2^30
    r[i] = (uint32_t)acc;
    acc >>= 32;                                                             // arithmetic
  }
  r[N] = (uint32_t)acc;
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = (r[i] >> kBingcdStep) | (r[i + 1] << (32 - kBingcdStep));
  return acc < 0;
}

template <int N>
B381_HD void bingcd_negate_if(uint32_t* x, bool neg) {
  const uint32_t mask = neg ? 0xFFFFFFFFu : 0u;
  uint64_t c = neg ? 1u : 0u;
#pragma unroll
  for (int i = 0; i < N; i++) {
    c += (uint64_t)(x[i] ^ mask);
    x[i] = (uint32_t)c;
    c >>= 32;
  }
}

// out = (u*f + v*g) / 2^30 mod m, for u, v in [0, m); out in [0, m).  minv32 = -m^-1 mod 2^32.
template <int N>
B381_HD void bingcd_lincomb_mod(const uint32_t* u, const uint32_t* v, int32_t f, int32_t g, const uint32_t* m, uint32_t minv32,
                                uint32_t* out) {
  uint32_t t[N + 1];
  int64_t acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    acc += (int64_t)(uint64_t)u[i] * f + (int64_t)(uint64_t)v[i] * g;
    t[i] = (uint32_t)acc;
    acc >>= 32;
  }
  // t + q m = 0 (mod 2^30) for q = t * (-m^-1) mod 2^30; |t| < 2^30 m, 0 <= q m < 2^30 m: the sum is in (-2^30 m, 2^31 m)
  const uint32_t q = (t[0] * minv32) & ((1u << kBingcdStep) - 1u);
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    c += (uint64_t)t[i] + (uint64_t)q * m[i];
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  const int64_t top = acc + (int64_t)c;            // signed word above the N limbs (|top| < 2^31: m < 2^(32N-1))
  t[N] = (uint32_t)top;
  uint32_t r[N];
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = (t[i] >> kBingcdStep) | (t[i + 1] << (32 - kBingcdStep));
  // the quotient lies in (-m, 2m): add m when negative, subtract m when >= m
  const bool neg = top < 0;
  uint32_t plus[N], minus[N];
  uint64_t ca = 0;
  int64_t br = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    ca += (uint64_t)r[i] + m[i];
    plus[i] = (uint32_t)ca;
    ca >>= 32;
    br += (int64_t)(uint64_t)r[i] - (int64_t)(uint64_t)m[i];
    minus[i] = (uint32_t)br;
    br >>= 32;
  }
  const bool ge = br >= 0;                         // r >= m
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = neg ? plus[i] : (ge ? minus[i] : r[i]);
}

// out = y^-1 mod m as integers (NOT a Montgomery-form operation), 0 for y = 0.  N 32-bit limbs, little endian.
// rounds = ceil((2 * bitlen(m) - 1) / 30) + 1.
template <int N>
B381_HD void bingcd_inverse(const uint32_t* y, const uint32_t* m, uint32_t minv32, int rounds, uint32_t* out) {
  uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
  for (int i = 0; i < N; i++) { a[i] = y[i]; b[i] = m[i]; u[i] = 0; v[i] = 0; }
  u[0] = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int round = 0; round < rounds; round++) {
    // ---- 64-bit approximations: low 30 bits + the 34 bits below the common top bit of (a, b); exact below 2^64
    int j = 0;
#pragma unroll
    for (int i = 1; i < N; i++) j = (a[i] | b[i]) != 0 ? i : j;
    uint32_t ha = 0, ma = 0, la = 0, hb = 0, mb = 0, lb = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      ha = i == j ? a[i] : ha;          hb = i == j ? b[i] : hb;
      ma = i == j - 1 ? a[i] : ma;      mb = i == j - 1 ? b[i] : mb;
      la = i == j - 2 ? a[i] : la;      lb = i == j - 2 ? b[i] : lb;
    }
    uint64_t abar = ((uint64_t)a[1] << 32) | a[0], bbar = ((uint64_t)b[1] << 32) | b[0];
    if (j >= 2) {
      const uint32_t hh = ha | hb;               // != 0
#if defined(__CUDA_ARCH__)
      const int s = __clz((int)hh);
#else
      const int s = __builtin_clz(hh);
#endif
      // top 64 bits of the 96-bit window (limbs j, j-1, j-2) once the common top bit is moved to bit 95
      const uint64_t wa = (((uint64_t)ha << 32) | ma), wb = (((uint64_t)hb << 32) | mb);
      const uint64_t ta = s ? (wa << s) | ((uint64_t)la >> (32 - s)) : wa;
      const uint64_t tb = s ? (wb << s) | ((uint64_t)lb >> (32 - s)) : wb;
      abar = ((ta >> kBingcdStep) << kBingcdStep) | (a[0] & ((1u << kBingcdStep) - 1u));
      bbar = ((tb >> kBingcdStep) << kBingcdStep) | (b[0] & ((1u << kBingcdStep) - 1u));
    }
    // ---- 30 binary-GCD steps on the approximations
    int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 6
    for (int i = 0; i < kBingcdStep; i++) {
      const uint64_t odd = 0 - (abar & 1);
      const uint64_t sw = odd & (abar < bbar ? ~0ull : 0ull);
      const uint64_t tx = (abar ^ bbar) & sw;
      abar ^= tx;
      bbar ^= tx;
      const int32_t tf = (f0 ^ f1) & (int32_t)sw, tg = (g0 ^ g1) & (int32_t)sw;
      f0 ^= tf; f1 ^= tf;
      g0 ^= tg; g1 ^= tg;
      abar -= bbar & odd;
      f0 -= f1 & (int32_t)odd;
      g0 -= g1 & (int32_t)odd;
      abar >>= 1;
      f1 <<= 1;
      g1 <<= 1;
    }
    // ---- apply the factors to (a, b) exactly and to (u, v) modulo m
    uint32_t na[N], nb[N];
    const bool nega = bingcd_lincomb<N>(a, b, f0, g0, na);
    const bool negb = bingcd_lincomb<N>(a, b, f1, g1, nb);
    bingcd_negate_if<N>(na, nega);
    bingcd_negate_if<N>(nb, negb);
    if (nega) { f0 = -f0; g0 = -g0; }
    if (negb) { f1 = -f1; g1 = -g1; }
    uint32_t nu[N], nv[N];
    bingcd_lincomb_mod<N>(u, v, f0, g0, m, minv32, nu);
    bingcd_lincomb_mod<N>(u, v, f1, g1, m, minv32, nv);
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
  }
  // gcd in b: 1 for every y != 0 (m prime); y = 0 leaves b = m and v = 0
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = v[i];
}

}  // namespace b381
