// BLS12-381 G1/G2 group law for the MSM hot path, templated on the coordinate field
// (fq_t for G1, fq2_t for G2).
//
// Working representation: XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity <=> ZZ == 0).
// The reference accumulates in Jacobian with a constant-time add that ALSO evaluates a full
// doubling every time (bls12-381/include/point.cuh:803-912, :885-886); here a bucket insertion
// is one 8M+2S mixed XYZZ addition and the exceptional cases (P = Q, P = -Q, identities) are
// handled by data-dependent branches -- MSM inputs are public (commitment bases, witness
// scalars of the prover itself), and the final affine result is representation independent,
// so bit-exactness is unaffected.
//
// Wire formats kept from the reference:
//   affine  (x, y) Montgomery, infinity = (0, 0)                  point.cuh:286-318
//   Jacobian (X, Y, Z), identity Z = 0                             point.cuh:455-525
#pragma once
#include "field.cuh"

namespace b381 {

template <class F> struct affine_t { F x, y; };
template <class F> struct jacobian_t { F x, y, z; };
template <class F> struct xyzz_t { F x, y, zz, zzz; };

using g1_affine = affine_t<fq_t>;
using g2_affine = affine_t<fq2_t>;
using g1_jac = jacobian_t<fq_t>;
using g2_jac = jacobian_t<fq2_t>;
using g1_xyzz = xyzz_t<fq_t>;
using g2_xyzz = xyzz_t<fq2_t>;

template <class F> B381_DI bool is_inf(const affine_t<F>& p) { return is_zero(p.x) && is_zero(p.y); }
template <class F> B381_DI bool is_inf(const xyzz_t<F>& p) { return is_zero(p.zz); }
template <class F> B381_DI xyzz_t<F> xyzz_identity() {
  return xyzz_t<F>{zero<F>(), zero<F>(), zero<F>(), zero<F>()};
}
template <class F> B381_DI xyzz_t<F> to_xyzz(const affine_t<F>& p) {
  if (is_inf(p)) return xyzz_identity<F>();
  return xyzz_t<F>{p.x, p.y, one<F>(), one<F>()};
}

// 2*(affine), never called with infinity.  mdbl-2008-s-1: 3M' (2 products + ...) see EFD.
template <class F> B381_DI xyzz_t<F> xyzz_dbl_affine(const affine_t<F>& p) {
  F u = dbl(p.y);
  F v = sqr(u);
  F w = mul(u, v);
  F s = mul(p.x, v);
  F x2 = sqr(p.x);
  F m = add(dbl(x2), x2);
  xyzz_t<F> r;
  r.x = sub(sqr(m), dbl(s));
  r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2*P in XYZZ (dbl-2008-s-1, a = 0).  y = 0 cannot occur on a prime-order subgroup point.
template <class F> B381_DI xyzz_t<F> xyzz_dbl(const xyzz_t<F>& p) {
  if (is_inf(p)) return p;
  F u = dbl(p.y);
  F v = sqr(u);
  F w = mul(u, v);
  F s = mul(p.x, v);
  F x2 = sqr(p.x);
  F m = add(dbl(x2), x2);
  xyzz_t<F> r;
  r.x = sub(sqr(m), dbl(s));
  r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
  r.zz = mul(v, p.zz);
  r.zzz = mul(w, p.zzz);
  return r;
}

// acc += q (affine), q != infinity.  madd-2008-s: 8M + 2S on the common path.
template <class F> B381_DI void xyzz_madd(xyzz_t<F>& acc, const affine_t<F>& q) {
  if (is_inf(acc)) {
    acc = xyzz_t<F>{q.x, q.y, one<F>(), one<F>()};
    return;
  }
  F u2 = mul(q.x, acc.zz);
  F s2 = mul(q.y, acc.zzz);
  F p = sub(u2, acc.x);
  F r = sub(s2, acc.y);
  if (is_zero(p)) {                 // same x: either doubling or cancellation (rare)
    if (is_zero(r)) acc = xyzz_dbl_affine(q);
    else acc = xyzz_identity<F>();
    return;
  }
  F pp = sqr(p);
  F ppp = mul(p, pp);
  F q1 = mul(acc.x, pp);
  F x3 = sub(sub(sqr(r), ppp), dbl(q1));
  F y3 = sub(mul(r, sub(q1, x3)), mul(acc.y, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = mul(acc.zz, pp);
  acc.zzz = mul(acc.zzz, ppp);
}

// acc += q (both XYZZ).  add-2008-s: 12M + 2S.
template <class F> B381_DI void xyzz_add(xyzz_t<F>& acc, const xyzz_t<F>& q) {
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc = q; return; }
  F u1 = mul(acc.x, q.zz);
  F u2 = mul(q.x, acc.zz);
  F s1 = mul(acc.y, q.zzz);
  F s2 = mul(q.y, acc.zzz);
  F p = sub(u2, u1);
  F r = sub(s2, s1);
  if (is_zero(p)) {
    if (is_zero(r)) acc = xyzz_dbl(acc);
    else acc = xyzz_identity<F>();
    return;
  }
  F pp = sqr(p);
  F ppp = mul(p, pp);
  F q1 = mul(u1, pp);
  F x3 = sub(sub(sqr(r), ppp), dbl(q1));
  F y3 = sub(mul(r, sub(q1, x3)), mul(s1, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = mul(mul(acc.zz, q.zz), pp);
  acc.zzz = mul(mul(acc.zzz, q.zzz), ppp);
}

template <class F> B381_DI affine_t<F> affine_neg(const affine_t<F>& p) { return affine_t<F>{p.x, neg(p.y)}; }
template <class F> B381_DI xyzz_t<F> xyzz_neg(const xyzz_t<F>& p) { return xyzz_t<F>{p.x, neg(p.y), p.zz, p.zzz}; }

// XYZZ -> affine Montgomery; identity -> (0,0).  One field inversion (of ZZZ):
//   1/ZZ = ZZ^2 * ... use  ZZ^3 = ZZZ^2  =>  1/ZZ = ZZ^2 / ZZZ^2 ... we simply invert ZZZ and
//   derive 1/ZZ = (1/ZZZ)^2 * ZZ^2 ... (ZZ^2/ZZZ^2 = ZZ^2/ZZ^3 = 1/ZZ).
template <class F> B381_DI affine_t<F> xyzz_to_affine(const xyzz_t<F>& p) {
  if (is_inf(p)) return affine_t<F>{zero<F>(), zero<F>()};
  F iz3 = inv_vartime(p.zzz);   // public data; ~6x shorter than a^(p-2) on a lone thread
  F t = mul(iz3, p.zz);      // ZZ/ZZZ
  F iz2 = sqr(t);            // ZZ^2/ZZZ^2 = 1/ZZ
  return affine_t<F>{mul(p.x, iz2), mul(p.y, iz3)};
}

// curve constant b (Montgomery): 4 on G1, 4(1 + u) on G2; y^2 = x^3 + b  (point.cuh:339-387)
template <class F> B381_DI F curve_b();
template <> B381_DI fq_t curve_b<fq_t>() { fq_t two = dbl(one<fq_t>()); return dbl(two); }
template <> B381_DI fq2_t curve_b<fq2_t>() { fq_t f = curve_b<fq_t>(); return fq2_t{f, f}; }
template <class F> B381_DI bool on_curve(const affine_t<F>& p) {
  return is_inf(p) || eq(sqr(p.y), add(mul(sqr(p.x), p.x), curve_b<F>()));
}

// Jacobian (reference wire format) -> XYZZ: ZZ = Z^2, ZZZ = Z^3
template <class F> B381_DI xyzz_t<F> jac_to_xyzz(const jacobian_t<F>& p) {
  if (is_zero(p.z)) return xyzz_identity<F>();
  F zz = sqr(p.z);
  return xyzz_t<F>{p.x, p.y, zz, mul(zz, p.z)};
}

}  // namespace b381
