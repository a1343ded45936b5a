/*
 * oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (csrc/, the C ABI, the Python host layer) never does.
 *
 * What is restated (file:line under the reference root):
 *   - Montgomery CIOS multiply / add / sub, 64-bit limbs + unsigned __int128:
 *       bls12-381/include/field.cuh:389-576            (field_add / field_sub / field_mul)
 *   - Fq2 Karatsuba / norm inverse:   bls12-381/include/point.cuh:131-225
 *   - Jacobian G1/G2 double / add / mixed add (a = 0):  bls12-381/include/point.cuh:610-912, :948-1259
 *   - affine (0,0) = infinity, Jacobian Z = 0 = identity: point.cuh:295-302, :469-486
 *   - signed-digit Pippenger, window heuristic get_optimal_c:
 *       bls12-381/include/msm.cuh:115-140, bls12-381/src/curve/msm_kernels.cu:69-143, :376-398, :578-596
 *   - ICICLE result convention (x, y, 1) standard form / (0, 1, 0): bls12-381/src/backend/icicle_curve_api.cu:134-229
 *   - NTT = best_fft contract (omega_k = ROOT_OF_UNITY^(2^(32-k)), forward natural->natural,
 *     inverse with omega^-1 then n^-1): core/ntt.rs:1488-1603; coset scaling include/ntt.cuh:123-183
 *
 * The arithmetic that the reference's MIDNIGHT_DEVICE=cpu path actually executes lives in
 * midnight-curves 0.2.0 (over blst) -- a path dependency `../curves` that is NOT under
 * /root/reference (Cargo.toml:19) and there is no Rust toolchain or libblst in this image, so this is
 * a "port" (cpu_baseline.kind = "port"), multi-threaded with OpenMP like rayon-backed
 * multi_exp / best_fft are.
 *
 * Pinning: tests/test_oracle.py checks this file against (a) the big-integer oracle oracle/pyref.py
 * on random inputs, (b) every known-answer the reference's tests hold for the path: constants
 * (tests/test_known_answer_vectors.cu:60-200), Fr 1*1=1, 0*1=0 (:221-236), 2*3=6, a*a^-1=1
 * (tests/test_field_properties.cu), 2P=P+P / O+P, 1*G=G, 0*G=O, sum i*G = 2080 G (core/msm.rs:1681-1694),
 * 5*G (:1667-1678), NTT(delta)=1..1 (core/ntt.rs:2059-2073), round trips, input 1..n at k=10
 * (tests/ntt_fft_comparison.rs:15-19).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------ moduli */
typedef struct { int n; u64 m[6]; u64 inv; u64 one[6]; u64 r2[6]; } modulus_t;

static const modulus_t FQ = {6,
  {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
  0x89f3fffcfffcfffdull,
  {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
  {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull}};
static const modulus_t FR = {4,
  {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull, 0, 0},
  0xfffffffeffffffffull,
  {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full, 0, 0},
  {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull, 0, 0}};
/* 7^((r-1)/2^32) in Montgomery form */
static const u64 FR_ROOT[4] = {0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull};

/* ------------------------------------------------------------------ generic n-limb Montgomery */
static inline int geq_n(const u64* a, const u64* b, int n) {
  for (int i = n - 1; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i];
  return 1;
}
static inline u64 addn(u64* r, const u64* a, const u64* b, int n) {
  u128 c = 0;
  for (int i = 0; i < n; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static inline u64 subn(u64* r, const u64* a, const u64* b, int n) {
  u64 br = 0;
  for (int i = 0; i < n; i++) { u128 d = (u128)a[i] - b[i] - br; r[i] = (u64)d; br = (u64)(d >> 64) & 1; }
  return br;
}
static inline void mod_add(u64* r, const u64* a, const u64* b, const modulus_t* M) {
  u64 t[6]; u64 c = addn(t, a, b, M->n);
  if (c || geq_n(t, M->m, M->n)) subn(t, t, M->m, M->n);
  memcpy(r, t, 8 * M->n);
}
static inline void mod_sub(u64* r, const u64* a, const u64* b, const modulus_t* M) {
  u64 t[6];
  if (subn(t, a, b, M->n)) addn(t, t, M->m, M->n);
  memcpy(r, t, 8 * M->n);
}
static inline void mod_neg(u64* r, const u64* a, const modulus_t* M) {
  u64 z[6] = {0, 0, 0, 0, 0, 0};
  mod_sub(r, z, a, M);
}
/* CIOS, field.cuh:510-576 */
static inline void mont_mul(u64* r, const u64* a, const u64* b, const modulus_t* M) {
  const int n = M->n;
  u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; i++) {
    u128 c = 0;
    for (int j = 0; j < n; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (u64)c; c >>= 64; }
    c += t[n]; t[n] = (u64)c; t[n + 1] = (u64)(c >> 64);
    u64 m = t[0] * M->inv;
    c = (u128)m * M->m[0] + t[0]; c >>= 64;
    for (int j = 1; j < n; j++) { c += (u128)m * M->m[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
    c += t[n]; t[n - 1] = (u64)c; t[n] = t[n + 1] + (u64)(c >> 64);
  }
  if (t[n] || geq_n(t, M->m, n)) subn(t, t, M->m, n);
  memcpy(r, t, 8 * n);
}
static inline int is_zero_n(const u64* a, int n) { u64 o = 0; for (int i = 0; i < n; i++) o |= a[i]; return o == 0; }
static void mod_pow(u64* r, const u64* a, const u64* e, int elimbs, const modulus_t* M) {
  u64 acc[6], base[6];
  memcpy(acc, M->one, 8 * M->n); memcpy(base, a, 8 * M->n);
  for (int i = 0; i < elimbs * 64; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) mont_mul(acc, acc, base, M);
    mont_mul(base, base, base, M);
  }
  memcpy(r, acc, 8 * M->n);
}
static void mod_inv(u64* r, const u64* a, const modulus_t* M) {   /* a^(m-2); inv(0)=0 (field.cuh:750-900) */
  u64 e[6]; memcpy(e, M->m, 8 * M->n); e[0] -= 2;
  mod_pow(r, a, e, M->n, M);
}

/* ------------------------------------------------------------------ exported field API (Montgomery in/out) */
void orc_fr_mul(const u64* a, const u64* b, u64* r) { mont_mul(r, a, b, &FR); }
void orc_fr_add(const u64* a, const u64* b, u64* r) { mod_add(r, a, b, &FR); }
void orc_fr_sub(const u64* a, const u64* b, u64* r) { mod_sub(r, a, b, &FR); }
void orc_fr_inv(const u64* a, u64* r) { mod_inv(r, a, &FR); }
void orc_fr_to_mont(const u64* a, u64* r) { mont_mul(r, a, FR.r2, &FR); }
void orc_fr_from_mont(const u64* a, u64* r) { u64 o[4] = {1, 0, 0, 0}; mont_mul(r, a, o, &FR); }
void orc_fq_mul(const u64* a, const u64* b, u64* r) { mont_mul(r, a, b, &FQ); }
void orc_fq_add(const u64* a, const u64* b, u64* r) { mod_add(r, a, b, &FQ); }
void orc_fq_sub(const u64* a, const u64* b, u64* r) { mod_sub(r, a, b, &FQ); }
void orc_fq_inv(const u64* a, u64* r) { mod_inv(r, a, &FQ); }
void orc_fq_to_mont(const u64* a, u64* r) { mont_mul(r, a, FQ.r2, &FQ); }
void orc_fq_from_mont(const u64* a, u64* r) { u64 o[6] = {1, 0, 0, 0, 0, 0}; mont_mul(r, a, o, &FQ); }
void orc_constants(u64* out /* fq: m,one,r2 (18) ; fr: m,one,r2 (12) ; invs (2) ; root (4) */) {
  memcpy(out, FQ.m, 48); memcpy(out + 6, FQ.one, 48); memcpy(out + 12, FQ.r2, 48);
  memcpy(out + 18, FR.m, 32); memcpy(out + 22, FR.one, 32); memcpy(out + 26, FR.r2, 32);
  out[30] = FQ.inv; out[31] = FR.inv; memcpy(out + 32, FR_ROOT, 32);
}
/* vecops on Montgomery vectors: op 0 add, 1 sub, 2 mul; a_scalar broadcasts a[0] (vec_ops.cu:63-118, :335-345) */
void orc_vecop(int op, int a_scalar, const u64* a, const u64* b, size_t n, u64* out) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) {
    const u64* x = a_scalar ? a : a + 4 * i;
    if (op == 0) mod_add(out + 4 * i, x, b + 4 * i, &FR);
    else if (op == 1) mod_sub(out + 4 * i, x, b + 4 * i, &FR);
    else mont_mul(out + 4 * i, x, b + 4 * i, &FR);
  }
}

/* ------------------------------------------------------------------ coordinate field abstraction: Fq (k=1) / Fq2 (k=2) */
typedef struct { u64 c[2][6]; } fe_t;   /* Fq uses c[0] only */

static inline void fe_add(fe_t* r, const fe_t* a, const fe_t* b, int k) { for (int i = 0; i < k; i++) mod_add(r->c[i], a->c[i], b->c[i], &FQ); }
static inline void fe_sub(fe_t* r, const fe_t* a, const fe_t* b, int k) { for (int i = 0; i < k; i++) mod_sub(r->c[i], a->c[i], b->c[i], &FQ); }
static inline void fe_neg(fe_t* r, const fe_t* a, int k) { for (int i = 0; i < k; i++) mod_neg(r->c[i], a->c[i], &FQ); }
static inline void fe_dbl(fe_t* r, const fe_t* a, int k) { fe_add(r, a, a, k); }
static inline int fe_is_zero(const fe_t* a, int k) { return is_zero_n(a->c[0], 6) && (k == 1 || is_zero_n(a->c[1], 6)); }
static inline void fe_zero(fe_t* r) { memset(r, 0, sizeof(*r)); }
static inline void fe_one(fe_t* r) { memset(r, 0, sizeof(*r)); memcpy(r->c[0], FQ.one, 48); }
static void fe_mul(fe_t* r, const fe_t* a, const fe_t* b, int k) {
  if (k == 1) { mont_mul(r->c[0], a->c[0], b->c[0], &FQ); return; }
  u64 v0[6], v1[6], s0[6], s1[6], s[6];                       /* point.cuh:147-164 (Karatsuba) */
  mont_mul(v0, a->c[0], b->c[0], &FQ); mont_mul(v1, a->c[1], b->c[1], &FQ);
  mod_add(s0, a->c[0], a->c[1], &FQ); mod_add(s1, b->c[0], b->c[1], &FQ);
  mont_mul(s, s0, s1, &FQ);
  mod_sub(s, s, v0, &FQ); mod_sub(s, s, v1, &FQ);
  mod_sub(r->c[0], v0, v1, &FQ); memcpy(r->c[1], s, 48);
}
static inline void fe_sqr(fe_t* r, const fe_t* a, int k) { fe_mul(r, a, a, k); }
static void fe_inv(fe_t* r, const fe_t* a, int k) {
  if (k == 1) { mod_inv(r->c[0], a->c[0], &FQ); return; }
  u64 n0[6], n1[6], n[6];                                     /* point.cuh:191-225 */
  mont_mul(n0, a->c[0], a->c[0], &FQ); mont_mul(n1, a->c[1], a->c[1], &FQ);
  mod_add(n, n0, n1, &FQ); mod_inv(n, n, &FQ);
  mont_mul(r->c[0], a->c[0], n, &FQ);
  mont_mul(n1, a->c[1], n, &FQ); mod_neg(r->c[1], n1, &FQ);
}
static void fe_from_mont(fe_t* r, const fe_t* a, int k) { u64 o[6] = {1, 0, 0, 0, 0, 0}; for (int i = 0; i < k; i++) mont_mul(r->c[i], a->c[i], o, &FQ); }

/* ------------------------------------------------------------------ Jacobian points */
typedef struct { fe_t x, y, z; } jac_t;
typedef struct { fe_t x, y; } aff_t;

static inline void jac_identity(jac_t* p) { fe_zero(&p->x); fe_one(&p->y); fe_zero(&p->z); }
static inline int jac_is_inf(const jac_t* p, int k) { return fe_is_zero(&p->z, k); }
static inline int aff_is_inf(const aff_t* p, int k) { return fe_is_zero(&p->x, k) && fe_is_zero(&p->y, k); }

static void jac_dbl(jac_t* r, const jac_t* p, int k) {          /* dbl-2009-l, a = 0 */
  if (jac_is_inf(p, k)) { *r = *p; return; }
  fe_t A, B, C, D, E, F, t;
  fe_sqr(&A, &p->x, k); fe_sqr(&B, &p->y, k); fe_sqr(&C, &B, k);
  fe_add(&t, &p->x, &B, k); fe_sqr(&t, &t, k); fe_sub(&t, &t, &A, k); fe_sub(&t, &t, &C, k); fe_dbl(&D, &t, k);
  fe_dbl(&E, &A, k); fe_add(&E, &E, &A, k);
  fe_sqr(&F, &E, k);
  fe_t x3, y3, z3;
  fe_dbl(&t, &D, k); fe_sub(&x3, &F, &t, k);
  fe_mul(&z3, &p->y, &p->z, k); fe_dbl(&z3, &z3, k);
  fe_sub(&t, &D, &x3, k); fe_mul(&y3, &E, &t, k);
  fe_dbl(&t, &C, k); fe_dbl(&t, &t, k); fe_dbl(&t, &t, k); fe_sub(&y3, &y3, &t, k);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(jac_t* r, const jac_t* p, const jac_t* q, int k) {  /* add-2007-bl with exceptions */
  if (jac_is_inf(p, k)) { *r = *q; return; }
  if (jac_is_inf(q, k)) { *r = *p; return; }
  fe_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
  fe_sqr(&z1z1, &p->z, k); fe_sqr(&z2z2, &q->z, k);
  fe_mul(&u1, &p->x, &z2z2, k); fe_mul(&u2, &q->x, &z1z1, k);
  fe_mul(&s1, &p->y, &q->z, k); fe_mul(&s1, &s1, &z2z2, k);
  fe_mul(&s2, &q->y, &p->z, k); fe_mul(&s2, &s2, &z1z1, k);
  fe_sub(&h, &u2, &u1, k); fe_sub(&rr, &s2, &s1, k);
  if (fe_is_zero(&h, k)) {
    if (fe_is_zero(&rr, k)) { jac_dbl(r, p, k); return; }
    jac_identity(r); return;
  }
  fe_dbl(&i, &h, k); fe_sqr(&i, &i, k); fe_mul(&j, &h, &i, k);
  fe_dbl(&rr, &rr, k); fe_mul(&v, &u1, &i, k);
  fe_t x3, y3, z3;
  fe_sqr(&x3, &rr, k); fe_sub(&x3, &x3, &j, k); fe_dbl(&t, &v, k); fe_sub(&x3, &x3, &t, k);
  fe_sub(&t, &v, &x3, k); fe_mul(&y3, &rr, &t, k); fe_mul(&t, &s1, &j, k); fe_dbl(&t, &t, k); fe_sub(&y3, &y3, &t, k);
  fe_add(&z3, &p->z, &q->z, k); fe_sqr(&z3, &z3, k); fe_sub(&z3, &z3, &z1z1, k); fe_sub(&z3, &z3, &z2z2, k); fe_mul(&z3, &z3, &h, k);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add_mixed(jac_t* r, const jac_t* p, const aff_t* q, int k) {
  if (aff_is_inf(q, k)) { *r = *p; return; }
  jac_t qq; qq.x = q->x; qq.y = q->y; fe_one(&qq.z);
  jac_add(r, p, &qq, k);
}
static void jac_to_affine(aff_t* r, const jac_t* p, int k) {     /* point.cuh:504-524 */
  if (jac_is_inf(p, k)) { fe_zero(&r->x); fe_zero(&r->y); return; }
  fe_t zi, zi2, zi3;
  fe_inv(&zi, &p->z, k); fe_sqr(&zi2, &zi, k); fe_mul(&zi3, &zi2, &zi, k);
  fe_mul(&r->x, &p->x, &zi2, k); fe_mul(&r->y, &p->y, &zi3, k);
}

/* limbs <-> structs: affine Montgomery, k*12 u64 */
static void aff_load(aff_t* p, const u64* src, int k) {
  memset(p, 0, sizeof(*p));
  for (int i = 0; i < k; i++) memcpy(p->x.c[i], src + 6 * i, 48);
  for (int i = 0; i < k; i++) memcpy(p->y.c[i], src + 6 * k + 6 * i, 48);
}
static void aff_store(u64* dst, const aff_t* p, int k) {
  for (int i = 0; i < k; i++) memcpy(dst + 6 * i, p->x.c[i], 48);
  for (int i = 0; i < k; i++) memcpy(dst + 6 * k + 6 * i, p->y.c[i], 48);
}
/* ICICLE result: (x, y, 1) standard form, identity (0, 1, 0); k*18 u64 */
static void result_store(u64* dst, const jac_t* p, int k) {
  aff_t a; jac_to_affine(&a, p, k);
  fe_t x, y, z;
  memset(dst, 0, 8 * 18 * k);
  if (jac_is_inf(p, k)) { dst[6 * k] = 1; return; }
  fe_from_mont(&x, &a.x, k); fe_from_mont(&y, &a.y, k);
  (void)z;
  for (int i = 0; i < k; i++) { memcpy(dst + 6 * i, x.c[i], 48); memcpy(dst + 6 * k + 6 * i, y.c[i], 48); }
  dst[12 * k] = 1;
}

/* ------------------------------------------------------------------ scalar multiplication, generators */
static const u64 G1X[6] = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull, 0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull};
static const u64 G1Y[6] = {0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull, 0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};
static const u64 G2XY[24] = {
  0xf5f28fa202940a10ull, 0xb3f5fb2687b4961aull, 0xa1a893b53e2ae580ull, 0x9894999d1a3caee9ull, 0x6f67b7631863366bull, 0x058191924350bcd7ull,
  0xa5a9c0759e23f606ull, 0xaaa0c59dbccd60c3ull, 0x3bb17e18e2867806ull, 0x1b1ab6cc8541b367ull, 0xc2b6ed0ef2158547ull, 0x11922a097360edf3ull,
  0x4c730af860494c4aull, 0x597cfa1f5e369c5aull, 0xe7e6856caa0a635aull, 0xbbefb5e96e0d495full, 0x07d3a975f0ef25a2ull, 0x0083fd8e7e80dae5ull,
  0xadc0fc92df64b05dull, 0x18aa270a2b1461dcull, 0x86adac6a3be4eba0ull, 0x79495c4ec93da33aull, 0xe7175850a43ccaedull, 0x0b2bc2a163de1bf2ull};

static void generator(aff_t* g, int k) {
  if (k == 1) { u64 t[12]; memcpy(t, G1X, 48); memcpy(t + 6, G1Y, 48); aff_load(g, t, 1); }
  else aff_load(g, G2XY, 2);
}
void orc_generator(int k, u64* out) { aff_t g; generator(&g, k); aff_store(out, &g, k); }

/* r = s * p, s = 4 standard-form limbs */
static void jac_mul(jac_t* r, const u64* s, const aff_t* p, int k) {
  jac_t acc; jac_identity(&acc);
  for (int i = 255; i >= 0; i--) {
    jac_dbl(&acc, &acc, k);
    if ((s[i >> 6] >> (i & 63)) & 1) jac_add_mixed(&acc, &acc, p, k);
  }
  *r = acc;
}
/* out = ICICLE result bytes of s*P ; P Montgomery affine, s standard form */
void orc_scalar_mul(int k, const u64* s, const u64* p, u64* out) {
  aff_t a; aff_load(&a, p, k); jac_t r; jac_mul(&r, s, &a, k); result_store(out, &r, k);
}
/* out[i] = (k0 + i*d) * G as Montgomery affine, i < n : series of known discrete logs */
void orc_gen_series(int k, const u64* k0, const u64* d, size_t n, u64* out) {
  aff_t g; generator(&g, k);
  jac_t D, P0; jac_mul(&D, d, &g, k); jac_mul(&P0, k0, &g, k);
  aff_t Da; jac_to_affine(&Da, &D, k);
  const size_t CH = 1024;
  size_t nch = (n + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1)
  for (long ch = 0; ch < (long)nch; ch++) {
    u64 off[4] = {(u64)ch * CH, 0, 0, 0};
    jac_t cur; jac_mul(&cur, off, &Da, k); jac_add(&cur, &cur, &P0, k);
    size_t beg = (size_t)ch * CH, end = beg + CH; if (end > n) end = n;
    size_t m = end - beg;
    jac_t* pts = (jac_t*)malloc(m * sizeof(jac_t));
    fe_t* pre = (fe_t*)malloc(m * sizeof(fe_t));
    fe_t acc; fe_one(&acc);
    for (size_t i = 0; i < m; i++) {       /* Montgomery batch inversion of the Z's (identity cannot occur) */
      pts[i] = cur; pre[i] = acc;
      if (!jac_is_inf(&cur, k)) fe_mul(&acc, &acc, &cur.z, k);
      jac_add_mixed(&cur, &cur, &Da, k);
    }
    fe_t inv; fe_inv(&inv, &acc, k);
    for (size_t i = m; i-- > 0;) {
      aff_t a;
      if (jac_is_inf(&pts[i], k)) { fe_zero(&a.x); fe_zero(&a.y); }
      else {
        fe_t zi, zi2, zi3; fe_mul(&zi, &inv, &pre[i], k); fe_mul(&inv, &inv, &pts[i].z, k);
        fe_sqr(&zi2, &zi, k); fe_mul(&zi3, &zi2, &zi, k);
        fe_mul(&a.x, &pts[i].x, &zi2, k); fe_mul(&a.y, &pts[i].y, &zi3, k);
      }
      aff_store(out + (size_t)12 * k * (beg + i), &a, k);
    }
    free(pts); free(pre);
  }
}
/* on-curve check of Montgomery affine points (point.cuh:339-387): y^2 = x^3 + b, b = 4 or 4(1+u) */
int orc_on_curve(int k, const u64* p) {
  aff_t a; aff_load(&a, p, k);
  if (aff_is_inf(&a, k)) return 1;
  fe_t y2, x3, b, four;
  fe_sqr(&y2, &a.y, k); fe_sqr(&x3, &a.x, k); fe_mul(&x3, &x3, &a.x, k);
  fe_one(&four); fe_dbl(&four, &four, 1); fe_dbl(&four, &four, 1);
  b = four; if (k == 2) memcpy(b.c[1], four.c[0], 48);
  fe_add(&x3, &x3, &b, k); fe_sub(&y2, &y2, &x3, k);
  return fe_is_zero(&y2, k);
}

/* ------------------------------------------------------------------ Pippenger MSM */
static int optimal_c(size_t n) {          /* include/msm.cuh:115-140 */
  int lg = 0; while (((size_t)1 << lg) < n) lg++;
  if (lg <= 8) return 7; if (lg <= 10) return 8; if (lg <= 12) return 10; if (lg <= 14) return 12;
  if (lg <= 16) return 13; if (lg <= 18) return 14; if (lg <= 20) return 15; return 16;
}
/* scalars: n x 4 limbs (Montgomery if scalars_mont); bases: Montgomery affine; out: ICICLE result (k*18 u64) */
int orc_msm(int k, const u64* scalars, int scalars_mont, const u64* bases, size_t n, int c, u64* out) {
  if (c <= 0) c = optimal_c(n ? n : 1);
  const int W = (256 + c - 1) / c;
  const size_t B = (size_t)1 << (c - 1);
  u64* sc = (u64*)malloc(n * 32 + 32);
  int32_t* digits = (int32_t*)malloc((n * W + 1) * sizeof(int32_t));
  if (!sc || !digits) return -1;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) {
    if (scalars_mont) orc_fr_from_mont(scalars + 4 * i, sc + 4 * i); else memcpy(sc + 4 * i, scalars + 4 * i, 32);
    /* signed digits, msm_kernels.cu:96-130 */
    u64 carry = 0;
    for (int w = 0; w < W; w++) {
      int bit = w * c, limb = bit >> 6, off = bit & 63;
      u64 d = 0;
      if (limb < 4) { d = sc[4 * i + limb] >> off; if (off + c > 64 && limb + 1 < 4) d |= sc[4 * i + limb + 1] << (64 - off); d &= ((u64)1 << c) - 1; }
      d += carry; carry = 0;
      int32_t sd = (int32_t)d;
      if (d > B) { sd = (int32_t)d - (1 << c); carry = 1; }
      digits[(size_t)w * n + i] = sd;
    }
  }
  jac_t* wsum = (jac_t*)malloc(W * sizeof(jac_t));
#pragma omp parallel for schedule(dynamic, 1)
  for (int w = 0; w < W; w++) {
    jac_t* bk = (jac_t*)malloc(B * sizeof(jac_t));
    for (size_t b = 0; b < B; b++) jac_identity(&bk[b]);
    for (size_t i = 0; i < n; i++) {
      int32_t d = digits[(size_t)w * n + i];
      if (!d) continue;
      aff_t p; aff_load(&p, bases + (size_t)12 * k * i, k);
      if (d < 0) { fe_neg(&p.y, &p.y, k); d = -d; }
      jac_add_mixed(&bk[d - 1], &bk[d - 1], &p, k);
    }
    jac_t run, tri; jac_identity(&run); jac_identity(&tri);      /* msm_kernels.cu:376-398 */
    for (long b = (long)B - 1; b >= 0; b--) { jac_add(&run, &run, &bk[b], k); jac_add(&tri, &tri, &run, k); }
    wsum[w] = tri;
    free(bk);
  }
  jac_t acc; jac_identity(&acc);                                   /* msm_kernels.cu:578-596 */
  for (int w = W - 1; w >= 0; w--) {
    for (int j = 0; j < c; j++) jac_dbl(&acc, &acc, k);
    jac_add(&acc, &acc, &wsum[w], k);
  }
  result_store(out, &acc, k);
  free(wsum); free(digits); free(sc);
  return 0;
}
/* sum_i s_i * k_i mod r  (standard-form inputs) -- the discrete-log side of the large-size check */
void orc_fr_dot(const u64* s, const u64* kk, size_t n, int s_mont, u64* out) {
  u64 acc[4] = {0, 0, 0, 0};
  for (size_t i = 0; i < n; i++) {
    u64 a[4], b[4], p[4];
    if (s_mont) memcpy(a, s + 4 * i, 32); else orc_fr_to_mont(s + 4 * i, a);
    orc_fr_to_mont(kk + 4 * i, b);
    mont_mul(p, a, b, &FR); mod_add(acc, acc, p, &FR);
  }
  orc_fr_from_mont(acc, out);
}

/* ------------------------------------------------------------------ NTT (best_fft contract) */
static void fr_pow_u64(u64* r, const u64* a, u64 e) { u64 ee[1] = {e}; mod_pow(r, a, ee, 1, &FR); }
void orc_omega(int log_n, u64* out) {           /* core/ntt.rs:1488-1494 */
  u64 w[4]; memcpy(w, FR_ROOT, 32);
  for (int i = log_n; i < 32; i++) mont_mul(w, w, w, &FR);
  memcpy(out, w, 32);
}
static size_t bitrev(size_t v, int bits) { size_t r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (v & 1); v >>= 1; } return r; }

/* in-place, natural in / natural out, Montgomery elements; inverse uses omega^-1 and n^-1 */
int orc_ntt(u64* a, int log_n, int inverse) {
  const size_t n = (size_t)1 << log_n;
  u64 w[4]; orc_omega(log_n, w);
  if (inverse) mod_inv(w, w, &FR);
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (i < j) { u64 t[4]; memcpy(t, a + 4 * i, 32); memcpy(a + 4 * i, a + 4 * j, 32); memcpy(a + 4 * j, t, 32); }
  }
  u64* tw = (u64*)malloc((n / 2 + 1) * 32);
  if (!tw) return -1;
  if (n >= 2) {
    const size_t CH = 4096, half = n / 2;
#pragma omp parallel for schedule(static)
    for (long c0 = 0; c0 < (long)half; c0 += CH) {
      u64 cur[4]; fr_pow_u64(cur, w, (u64)c0);
      for (size_t j = c0; j < (size_t)c0 + CH && j < half; j++) { memcpy(tw + 4 * j, cur, 32); mont_mul(cur, cur, w, &FR); }
    }
  }
  for (int s = 0; s < log_n; s++) {
    const size_t m = (size_t)1 << s, stride = n >> (s + 1);
#pragma omp parallel for schedule(static)
    for (long b = 0; b < (long)(n / 2); b++) {
      size_t grp = (size_t)b >> s, j = (size_t)b & (m - 1);
      u64* u = a + 4 * (grp * 2 * m + j);
      u64* v = u + 4 * m;
      u64 t[4]; mont_mul(t, v, tw + 4 * (j * stride), &FR);
      mod_sub(v, u, t, &FR); mod_add(u, u, t, &FR);
    }
  }
  if (inverse) {
    u64 nn[4] = {n, 0, 0, 0}, ninv[4];
    orc_fr_to_mont(nn, nn); mod_inv(ninv, nn, &FR);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) mont_mul(a + 4 * i, a + 4 * i, ninv, &FR);
  }
  free(tw);
  return 0;
}
/* coset: forward x[i]*g^i then NTT; inverse iNTT then *g^-i  (include/ntt.cuh:123-183); g Montgomery */
int orc_coset_ntt(u64* a, int log_n, int inverse, const u64* g) {
  const size_t n = (size_t)1 << log_n;
  u64 gg[4]; memcpy(gg, g, 32);
  if (inverse) { if (orc_ntt(a, log_n, 1)) return -1; mod_inv(gg, gg, &FR); }
  u64 cur[4]; memcpy(cur, FR.one, 32);
  for (size_t i = 0; i < n; i++) { mont_mul(a + 4 * i, a + 4 * i, cur, &FR); mont_mul(cur, cur, gg, &FR); }
  if (!inverse) return orc_ntt(a, log_n, 0);
  return 0;
}
/* sum_j a[j] z^j by Horner over per-thread chunks (a, z, out Montgomery).  Shares no code with orc_ntt: the tests
 * use it to re-evaluate single NTT outputs y[i] = A(omega^i) at sizes where a Python loop is out of reach. */
void orc_fr_poly_eval(const u64* a, size_t n, const u64* z, u64* out) {
  const size_t CH = (size_t)1 << 16;
  const size_t chunks = (n + CH - 1) / CH;
  u64* part = (u64*)malloc(chunks * 32);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < (long)chunks; c++) {
    size_t lo = (size_t)c * CH, hi = lo + CH < n ? lo + CH : n;
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t j = hi; j-- > lo;) { mont_mul(acc, acc, z, &FR); mod_add(acc, acc, a + 4 * j, &FR); }
    memcpy(part + 4 * c, acc, 32);
  }
  u64 zc[4], acc[4] = {0, 0, 0, 0};
  fr_pow_u64(zc, z, (u64)CH);
  for (size_t c = chunks; c-- > 0;) { mont_mul(acc, acc, zc, &FR); mod_add(acc, acc, part + 4 * c, &FR); }
  memcpy(out, acc, 32);
  free(part);
}
void orc_bit_reverse(u64* a, int log_n) {
  const size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (i < j) { u64 t[4]; memcpy(t, a + 4 * i, 32); memcpy(a + 4 * i, a + 4 * j, 32); memcpy(a + 4 * j, t, 32); }
  }
}
/* torchrun exports OMP_NUM_THREADS=1 to its workers; bench.py --impl reference asks for every core explicitly */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* SplitMix64 -> canonical Fr, same acceptance rule as tests/security_audit_tests.cuh:400-416 */
void orc_random_fr(u64 seed, size_t n, u64* out) {
  u64 s = seed;
  for (size_t i = 0; i < n; i++) {
    for (;;) {
      u64 l[4];
      for (int j = 0; j < 4; j++) {
        s += 0x9E3779B97F4A7C15ull; u64 z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; l[j] = z ^ (z >> 31);
      }
      l[3] &= 0x7fffffffffffffffull;
      if (!geq_n(l, FR.m, 4)) { memcpy(out + 4 * i, l, 32); break; }
    }
  }
}
