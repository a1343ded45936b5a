// Endomorphisms of BLS12-381 used off the MSM path: per-thread bodies, host-testable.
//
//   G1  phi(x, y) = (beta x, y), beta^3 = 1 in Fq: multiplication by lambda = z^2 - 1 on the order-r subgroup.
//       r = lambda^2 + lambda + 1 EXACTLY, so a scalar k < r splits by plain Euclidean division,
//         k = k1 + k2 lambda,  k2 = floor(k / lambda) <= lambda + 1,  k1 = k mod lambda,  both < 2^128, both >= 0
//       and k P = k1 P + k2 phi(P).  The same identity gives the membership test: for ANY P on E(Fq),
//       phi(P) = [lambda] P  implies  [lambda^2 + lambda + 1] P = [r] P = O (phi^2 + phi + 1 = 0 on the curve), and
//       E(Fq) has a single subgroup of order r; the converse holds by the choice of beta.
//   G2  psi = twist o Frobenius o untwist: multiplication by z on the order-r subgroup; P in G2 <=> psi(P) = [z] P
//       (M. Scott, "A note on group membership tests for G1, G2 and GT on BLS pairing-friendly curves", 2021).
//
// Replaces: the GLV constants / g1_endomorphism / glv_decompose of bls12-381/src/curve/point_ops.cu:103-320 (the
// reference approximates k2 by k >> 128 and carries a sign; `GLV_ENABLED` is off by default there) and the two
// subgroup checks the reference leaves as TODO (include/point.cuh:419-448).
#pragma once
#include "curve.cuh"

namespace b381 {

B381_DI fq_t glv_beta() { return fq_t{GLV_BETA_MONT_INIT}; }

// k (canonical, < r) -> k1 + k2 * lambda, each two 64-bit limbs
B381_DI void glv_decompose(const fr_t& k, uint64_t k1[2], uint64_t k2[2]) {
  const uint64_t lam[2] = GLV_LAMBDA_INIT;
  const uint64_t g[3] = GLV_RECIP_INIT;
  // q = floor(k * g / 2^256): limbs 4.. of the 7-limb product; never above floor(k / lambda), at most 2 below
  uint64_t prod[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    unsigned __int128 carry = 0;
    for (int j = 0; j < 3; j++) {
      unsigned __int128 t = (unsigned __int128)k.l[i] * g[j] + prod[i + j] + carry;
      prod[i + j] = (uint64_t)t;
      carry = t >> 64;
    }
    prod[i + 3] = (uint64_t)carry;
  }
  uint64_t q[3] = {prod[4], prod[5], prod[6]};
  // rem = k - q * lambda, kept modulo 2^192 (the true value is below 3 lambda < 2^130)
  uint64_t ql[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) {
    unsigned __int128 carry = 0;
    for (int j = 0; j < 2 && i + j < 3; j++) {
      unsigned __int128 t = (unsigned __int128)q[i] * lam[j] + ql[i + j] + carry;
      ql[i + j] = (uint64_t)t;
      carry = t >> 64;
    }
    if (i == 0) ql[2] = (uint64_t)carry;          // row 0 carries into limb 2; later rows only matter modulo 2^192
  }
  uint64_t rem[3];
  {
    uint64_t borrow = 0;
    for (int i = 0; i < 3; i++) {
      unsigned __int128 t = (unsigned __int128)k.l[i] - ql[i] - borrow;
      rem[i] = (uint64_t)t;
      borrow = (uint64_t)(t >> 64) ? 1 : 0;
    }
  }
  for (int it = 0; it < 3; it++) {
    const bool ge = rem[2] != 0 || rem[1] > lam[1] || (rem[1] == lam[1] && rem[0] >= lam[0]);
    if (!ge) break;
    unsigned __int128 t = (unsigned __int128)rem[0] - lam[0];
    rem[0] = (uint64_t)t;
    const uint64_t b = (uint64_t)(t >> 64) ? 1 : 0;
    t = (unsigned __int128)rem[1] - lam[1] - b;
    rem[1] = (uint64_t)t;
    rem[2] -= (uint64_t)(t >> 64) ? 1 : 0;
    if (++q[0] == 0 && ++q[1] == 0) ++q[2];
  }
  k1[0] = rem[0]; k1[1] = rem[1];
  k2[0] = q[0]; k2[1] = q[1];      // q <= lambda + 1 < 2^128
}

// [e] P for a 64-bit e, left-to-right double-and-add on XYZZ (|z| has Hamming weight 6: 63 doublings, 5 additions)
template <class F>
B381_DI xyzz_t<F> xyzz_mul_u64(const xyzz_t<F>& p, uint64_t e) {
  xyzz_t<F> r = xyzz_identity<F>();
  if (e == 0 || is_inf(p)) return r;
  int top = 63;
  while (!((e >> top) & 1)) top--;
#ifndef B381_HOST_TEST
#pragma unroll 1
#endif
  for (int i = top; i >= 0; i--) {
    r = xyzz_dbl(r);
    if ((e >> i) & 1) xyzz_add(r, p);
  }
  return r;
}

// does the XYZZ point q equal the affine point (x, y)?  (neither at infinity)
template <class F>
B381_DI bool xyzz_eq_affine(const xyzz_t<F>& q, const F& x, const F& y) {
  return eq(q.x, mul(x, q.zz)) && eq(q.y, mul(y, q.zzz));
}

// P on E(Fq) (not checked here) lies in G1?  phi(P) == [lambda] P, [lambda] P = [|z|]([|z|] P) - P
B381_DI bool g1_in_subgroup(const affine_t<fq_t>& p) {
  if (is_inf(p)) return true;
  xyzz_t<fq_t> q = xyzz_mul_u64(xyzz_mul_u64(to_xyzz(p), BLS_Z_ABS), BLS_Z_ABS);   // [z^2] P
  xyzz_madd(q, affine_neg(p));                                                       // [z^2 - 1] P
  if (is_inf(q)) return false;
  return xyzz_eq_affine(q, mul(p.x, glv_beta()), p.y);
}

B381_DI fq2_t fq2_conj(const fq2_t& a) { return fq2_t{a.c0, neg(a.c1)}; }

// P on E'(Fq2) lies in G2?  psi(P) == [z] P = -[|z|] P
B381_DI bool g2_in_subgroup(const affine_t<fq2_t>& p) {
  if (is_inf(p)) return true;
  const fq2_t cx = PSI_CX_MONT_INIT, cy = PSI_CY_MONT_INIT;
  const fq2_t px = mul(fq2_conj(p.x), cx), py = mul(fq2_conj(p.y), cy);
  const xyzz_t<fq2_t> q = xyzz_mul_u64(to_xyzz(p), BLS_Z_ABS);
  if (is_inf(q)) return false;
  return xyzz_eq_affine(q, px, neg(py));
}

// k P by GLV: 4-bit windows over k1 and k2 share ONE table of multiples of P, because phi commutes with the group
// law: phi(j P) = (beta X, Y, ZZ, ZZZ).  128 doublings + <= 64 additions instead of 252 + <= 64.
// T: 15 XYZZ entries, T[j] = (j + 1) P.
B381_DI void g1_mul_table(const affine_t<fq_t>& p, xyzz_t<fq_t>* T) {
  T[0] = to_xyzz(p);
  for (int j = 1; j < 15; j++) {
    if (j & 1) T[j] = xyzz_dbl(T[(j - 1) >> 1]);      // (j + 1) even: 2 * ((j + 1) / 2) P
    else { T[j] = T[j - 1]; xyzz_madd(T[j], p); }
  }
}

// k mod r for any 256-bit k (at most two subtractions: 2^256 < 3r); the split below needs k < r
B381_DI fr_t fr_canonical(fr_t k) {
  const uint64_t R[4] = FR_MODULUS_INIT;
  for (int it = 0; it < 2; it++) {
    uint64_t d[4], borrow = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 t = (unsigned __int128)k.l[i] - R[i] - borrow;
      d[i] = (uint64_t)t;
      borrow = (uint64_t)(t >> 64) ? 1 : 0;
    }
    if (!borrow) for (int i = 0; i < 4; i++) k.l[i] = d[i];
  }
  return k;
}

B381_DI xyzz_t<fq_t> g1_mul_glv(const affine_t<fq_t>& p, const fr_t& k_in) {
  xyzz_t<fq_t> acc = xyzz_identity<fq_t>();
  if (is_inf(p)) return acc;
  const fr_t k = fr_canonical(k_in);
  uint64_t k1[2], k2[2];
  glv_decompose(k, k1, k2);
  xyzz_t<fq_t> T[15];
  g1_mul_table(p, T);
  const fq_t beta = glv_beta();
#ifndef B381_HOST_TEST
#pragma unroll 1
#endif
  for (int w = 31; w >= 0; w--) {
    for (int d = 0; d < 4; d++) acc = xyzz_dbl(acc);
    const uint32_t d1 = (uint32_t)(k1[w >> 4] >> ((w & 15) * 4)) & 15u;
    const uint32_t d2 = (uint32_t)(k2[w >> 4] >> ((w & 15) * 4)) & 15u;
    if (d1) xyzz_add(acc, T[d1 - 1]);
    if (d2) {
      xyzz_t<fq_t> q = T[d2 - 1];
      q.x = mul(q.x, beta);
      xyzz_add(acc, q);
    }
  }
  return acc;
}

// the plain windowed method over all 256 bits (bls12_381_g1_scalar_mul, point_ops.cu:480-547): comparator of the above
B381_DI xyzz_t<fq_t> g1_mul_window(const affine_t<fq_t>& p, const fr_t& k) {
  xyzz_t<fq_t> acc = xyzz_identity<fq_t>();
  if (is_inf(p)) return acc;
  xyzz_t<fq_t> T[15];
  g1_mul_table(p, T);
#ifndef B381_HOST_TEST
#pragma unroll 1
#endif
  for (int w = 63; w >= 0; w--) {
    for (int d = 0; d < 4; d++) acc = xyzz_dbl(acc);
    const uint32_t dg = (uint32_t)(k.l[w >> 4] >> ((w & 15) * 4)) & 15u;
    if (dg) xyzz_add(acc, T[dg - 1]);
  }
  return acc;
}

}  // namespace b381
