// Lazy-reduced, carry-free Fq for the G1 bucket-accumulation loop: wrappers over the generated
// routines in fq_unsat.cuh (13 x 30-bit limbs, Montgomery radix 2^390) and the XYZZ mixed addition
// written against them.  Design, bounds and the reason (IMAD.WIDE.X issues at half the rate of plain
// IMAD.WIDE on B200) are in csrc/gen/gen_unsat.py.
//
// Value bounds (multiples of p) that make every routine's precondition hold inside the loop:
//   product / square output            < 1.2       (inputs up to 64 x 10)
//   X1 < 7.2, Y1 < 3.2, ZZ1, ZZZ1 < 1.2            (accumulator between insertions)
//   P = U2 - X1 + 8p < 9.2,  R = S2 - Y1 + 4p < 5.2,  Q - X3 + 8p < 9.2
// all far below the 2^390 / p = 630 that 13 limbs can hold.
#pragma once
#ifndef B381_HOST_TEST
#include "fq_unsat.cuh"
#endif
#include "curve.cuh"

namespace b381 {

#ifndef B381_HOST_TEST
struct fqu_t { uint32_t v[13]; };

// The product and the square are real (non-inlined) functions: the mixed addition uses them ten
// times and, inlined, its 165 KB of straight-line code thrashed the instruction cache
// (sm__icc_request_hit_rate 70 %, "no_instruction" the top stall; profiles/r01_lazy_inlined_icache.txt).
// ptxas passes and returns the 13-limb structs entirely in registers (no stack traffic).
static __device__ __noinline__ fqu_t fqu_mul(fqu_t a, fqu_t b) { fqu_t r; fqu_mul_raw(r.v, a.v, b.v); return r; }
static __device__ __noinline__ fqu_t fqu_sqr(fqu_t a) { fqu_t r; fqu_sqr_raw(r.v, a.v); return r; }
B381_DI fqu_t fqu_sub_k2(const fqu_t& a, const fqu_t& b) { fqu_t r; fqu_sub_k2_raw(r.v, a.v, b.v); return r; }
B381_DI fqu_t fqu_sub_k4(const fqu_t& a, const fqu_t& b) { fqu_t r; fqu_sub_k4_raw(r.v, a.v, b.v); return r; }
B381_DI fqu_t fqu_sub_k8(const fqu_t& a, const fqu_t& b) { fqu_t r; fqu_sub_k8_raw(r.v, a.v, b.v); return r; }
B381_DI fqu_t fqu_sub2_k4(const fqu_t& a, const fqu_t& b) { fqu_t r; fqu_sub2_k4_raw(r.v, a.v, b.v); return r; }
B381_DI fqu_t fqu_neg_k2(const fqu_t& a) { fqu_t r; fqu_neg_k2_raw(r.v, a.v); return r; }
B381_DI fqu_t fqu_one() { return fqu_t{FQU_ONE_INIT}; }

// wire Montgomery form (R = 2^384, canonical) -> internal: value * 2^6, no reduction needed
B381_DI fqu_t fqu_from_wire(const fq_t& a) {
  uint32_t w[12];
#pragma unroll
  for (int i = 0; i < 6; i++) { w[2 * i] = (uint32_t)a.l[i]; w[2 * i + 1] = (uint32_t)(a.l[i] >> 32); }
  fqu_t r;
  fqu_from_wire_raw(r.v, w);
  return r;
}
// internal -> canonical wire Montgomery form: one product with 2^384, one conditional subtraction
B381_DI fq_t fqu_to_wire(const fqu_t& a) {
  const fqu_t k = {FQU_TO_WIRE_INIT};
  fqu_t t = fqu_mul(a, k);
  uint32_t w[12];
  fqu_pack_canonical_raw(w, t.v);
  fq_t r;
#pragma unroll
  for (int i = 0; i < 6; i++) r.l[i] = ((uint64_t)w[2 * i + 1] << 32) | w[2 * i];
  return r;
}
// value == 0 mod p, for a product output (< 2p, normalised)
B381_DI bool fqu_is_zero_mod_p(const fqu_t& a) {
  uint32_t z[2];
  fqu_zero_test_raw(z, a.v);
  return z[0] == 0 || z[1] == 0;
}
#else
// ---- CPU-only unit-test stand-in (tests/host): same interface, canonical values underneath ----
struct fqu_t { fq_t c; };
B381_DI fqu_t fqu_mul(const fqu_t& a, const fqu_t& b) { return fqu_t{mul(a.c, b.c)}; }
B381_DI fqu_t fqu_sqr(const fqu_t& a) { return fqu_t{sqr(a.c)}; }
B381_DI fqu_t fqu_sub_k2(const fqu_t& a, const fqu_t& b) { return fqu_t{sub(a.c, b.c)}; }
B381_DI fqu_t fqu_sub_k4(const fqu_t& a, const fqu_t& b) { return fqu_t{sub(a.c, b.c)}; }
B381_DI fqu_t fqu_sub_k8(const fqu_t& a, const fqu_t& b) { return fqu_t{sub(a.c, b.c)}; }
B381_DI fqu_t fqu_sub2_k4(const fqu_t& a, const fqu_t& b) { return fqu_t{sub(a.c, dbl(b.c))}; }
B381_DI fqu_t fqu_neg_k2(const fqu_t& a) { return fqu_t{neg(a.c)}; }
B381_DI fqu_t fqu_one() { return fqu_t{one<fq_t>()}; }
B381_DI fqu_t fqu_from_wire(const fq_t& a) { return fqu_t{a}; }
B381_DI fq_t fqu_to_wire(const fqu_t& a) { return a.c; }
B381_DI bool fqu_is_zero_mod_p(const fqu_t& a) { return is_zero(a.c); }
#endif

// XYZZ accumulator in the lazy domain; `inf` replaces the ZZ == 0 test of xyzz_t
struct g1_lazy_acc {
  fqu_t x, y, zz, zzz;
  bool inf;
};

B381_DI void lazy_set_affine(g1_lazy_acc& acc, const fq_t& x, const fq_t& y) {
  const fqu_t one_i = fqu_one();
  // x64 values are reduced once (multiply by the internal one) so the loop bounds hold from the start
  acc.x = fqu_mul(fqu_from_wire(x), one_i);
  acc.y = fqu_mul(fqu_from_wire(y), one_i);
  acc.zz = one_i;
  acc.zzz = one_i;
  acc.inf = false;
}

// acc += (qx, qy) (affine, wire Montgomery form, NOT infinity).  madd-2008-s, 8M + 2S.
B381_DI void lazy_madd(g1_lazy_acc& acc, const fq_t& qx, const fq_t& qy) {
  if (acc.inf) { lazy_set_affine(acc, qx, qy); return; }
  fqu_t u2 = fqu_mul(fqu_from_wire(qx), acc.zz);
  fqu_t s2 = fqu_mul(fqu_from_wire(qy), acc.zzz);
  fqu_t p = fqu_sub_k8(u2, acc.x);
  fqu_t r = fqu_sub_k4(s2, acc.y);
  fqu_t pp = fqu_sqr(p);
  if (fqu_is_zero_mod_p(pp)) {                 // same x: doubling or cancellation (rare)
    if (fqu_is_zero_mod_p(fqu_sqr(r))) {
      xyzz_t<fq_t> d = xyzz_dbl_affine(affine_t<fq_t>{qx, qy});   // saturated path, once in a blue moon
      const fqu_t one_i = fqu_one();
      acc.x = fqu_mul(fqu_from_wire(d.x), one_i);
      acc.y = fqu_mul(fqu_from_wire(d.y), one_i);
      acc.zz = fqu_mul(fqu_from_wire(d.zz), one_i);
      acc.zzz = fqu_mul(fqu_from_wire(d.zzz), one_i);
    } else {
      acc.inf = true;
    }
    return;
  }
  fqu_t ppp = fqu_mul(p, pp);
  fqu_t q = fqu_mul(acc.x, pp);
  fqu_t x3 = fqu_sub2_k4(fqu_sub_k2(fqu_sqr(r), ppp), q);        // R^2 - PPP - 2Q
  fqu_t y3 = fqu_sub_k2(fqu_mul(r, fqu_sub_k8(q, x3)), fqu_mul(acc.y, ppp));
  acc.zz = fqu_mul(acc.zz, pp);
  acc.zzz = fqu_mul(acc.zzz, ppp);
  acc.x = x3;
  acc.y = y3;
}

B381_DI xyzz_t<fq_t> lazy_to_xyzz(const g1_lazy_acc& acc) {
  if (acc.inf) return xyzz_identity<fq_t>();
  return xyzz_t<fq_t>{fqu_to_wire(acc.x), fqu_to_wire(acc.y), fqu_to_wire(acc.zz), fqu_to_wire(acc.zzz)};
}

}  // namespace b381
