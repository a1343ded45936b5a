// Fr vector operations (Montgomery domain), 128-bit coalesced accesses: the five registered element-wise ops
// plus the unregistered ones the reference's callers use on the same data: vector sum,
// batched inversion, bit-reversal permutation, Montgomery conversion.
// Replaces vec_{add,sub,mul}_kernel / scalar_vec_{mul,add}_kernel (bls12-381/src/field/vec_ops.cu:63-118,
// :335-345) and the run_vec_op staging helper (bls12-381/src/backend/icicle_field_api.cu:133-334).
// HBM-bound: 96 B/element (two reads + one write), 64 B/element for scalar (op) vector.
#include <cstring>

#include "common.cuh"
#include "field.cuh"
#include "fr_inv.cuh"

namespace b381 {

enum VecOp { VADD = 0, VSUB = 1, VMUL = 2 };

B381_DI fr_t ld_fr(const fr_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fr_t r;
  r.l[0] = ((uint64_t)a.y << 32) | a.x; r.l[1] = ((uint64_t)a.w << 32) | a.z;
  r.l[2] = ((uint64_t)b.y << 32) | b.x; r.l[3] = ((uint64_t)b.w << 32) | b.z;
  return r;
}
B381_DI void st_fr(fr_t* p, const fr_t& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4((uint32_t)v.l[0], (uint32_t)(v.l[0] >> 32), (uint32_t)v.l[1], (uint32_t)(v.l[1] >> 32));
  q[1] = make_uint4((uint32_t)v.l[2], (uint32_t)(v.l[2] >> 32), (uint32_t)v.l[3], (uint32_t)(v.l[3] >> 32));
}

template <int OP> B381_DI fr_t apply(const fr_t& a, const fr_t& b) {
  if (OP == VADD) return add(a, b);
  if (OP == VSUB) return sub(a, b);
  return mul(a, b);
}

// grid-stride; a_is_scalar broadcasts a[0]
template <int OP, bool A_SCALAR>
__global__ void __launch_bounds__(256) k_vecop(const fr_t* a, const fr_t* b, uint64_t n, fr_t* out) {
  fr_t s;
  if (A_SCALAR) s = ld_fr(a);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    fr_t x = A_SCALAR ? s : ld_fr(a + i);
    fr_t y = ld_fr(b + i);
    st_fr(out + i, apply<OP>(x, y));
  }
}

template <int OP, bool A_SCALAR>
static int vecop_entry(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* cfg, b381_fr* out) {
  if (!cfg) return B381_INVALID_POINTER;
  if (n == 0) return B381_SUCCESS;
  if (!a || !b || !out) return B381_INVALID_POINTER;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  TraceRange trace(OP == VADD ? "b381_vec add" : OP == VSUB ? "b381_vec sub" : "b381_vec mul", (long long)n);
  {
    Scratch sc(st);
    const fr_t *da, *db;
    if ((e = stage_in(sc, (const fr_t*)a, A_SCALAR ? 1 : n, cfg->is_a_on_device, &da)) != cudaSuccess) return map_cuda_error(e);
    if ((e = stage_in(sc, (const fr_t*)b, n, cfg->is_b_on_device, &db)) != cudaSuccess) return map_cuda_error(e);
    fr_t* dout = (fr_t*)out;
    if (!cfg->is_result_on_device && (e = sc.alloc(&dout, n)) != cudaSuccess) return map_cuda_error(e);
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;   // 148 SMs x 8 resident CTAs x 4 waves of grid-stride
    k_vecop<OP, A_SCALAR><<<(unsigned)blocks, 256, 0, st>>>(da, db, n, dout);
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->is_result_on_device) {
      e = cudaMemcpyAsync(out, dout, n * sizeof(fr_t), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return map_cuda_error(e);
    }
  }
  if (!cfg->is_async) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return map_cuda_error(e);
  }
  return B381_SUCCESS;
}

// Upstream ICICLE v4 batch semantics of scalar (op) vector (VecOpsConfig.batch_size / columns_batch, which the reference's
// struct lacks but its Rust layer sets, core/vecops.rs:345-346): `batch` scalars, `batch` vectors of n elements stored
// one after the other (rows) or interleaved (columns_batch: element j of vector k at j * batch + k).
template <int OP>
__global__ void __launch_bounds__(256) k_scalar_vec_batch(const fr_t* scalars, const fr_t* vec, uint64_t n, uint32_t batch,
                                                          bool columns, uint64_t total, fr_t* out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = columns ? i % batch : i / n;
    st_fr(out + i, apply<OP>(ld_fr(scalars + k), ld_fr(vec + i)));
  }
}

template <int OP>
static int scalar_batch_entry(const b381_fr* a, const b381_fr* b, uint64_t n, int batch_size, bool columns,
                              const b381_vecops_config* cfg, b381_fr* out) {
  if (!cfg) return B381_INVALID_POINTER;
  if (batch_size < 1) return B381_INVALID_ARGUMENT;
  if (batch_size == 1) return vecop_entry<OP, true>(a, b, n, cfg, out);
  if (n == 0) return B381_SUCCESS;
  if (!a || !b || !out) return B381_INVALID_POINTER;
  const uint64_t total = n * (uint64_t)batch_size;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  {
    Scratch sc(st);
    const fr_t *da, *db;
    if ((e = stage_in(sc, (const fr_t*)a, (size_t)batch_size, cfg->is_a_on_device, &da)) != cudaSuccess) return map_cuda_error(e);
    if ((e = stage_in(sc, (const fr_t*)b, total, cfg->is_b_on_device, &db)) != cudaSuccess) return map_cuda_error(e);
    fr_t* dout = (fr_t*)out;
    if (!cfg->is_result_on_device && (e = sc.alloc(&dout, total)) != cudaSuccess) return map_cuda_error(e);
    uint64_t blocks = (total + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    k_scalar_vec_batch<OP><<<(unsigned)blocks, 256, 0, st>>>(da, db, n, (uint32_t)batch_size, columns, total, dout);
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->is_result_on_device && (e = cudaMemcpyAsync(out, dout, total * sizeof(fr_t), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
      return map_cuda_error(e);
  }
  if (!cfg->is_async && (e = cudaStreamSynchronize(st)) != cudaSuccess) return map_cuda_error(e);
  return B381_SUCCESS;
}

// ---- unregistered-but-needed ops -------------------------------------------------------
// out = sum a[i]: grid-stride partial per thread, shuffle + shared-memory tree per CTA, second launch folds the
// CTA partials.  Modular addition is associative, so the result is bit-exact whatever the order.
// (reference: vec_sum_partial_kernel / vec_sum_cuda, bls12-381/src/field/vec_ops.cu:350-385, :479-520)
B381_DI fr_t shfl_down_fr(const fr_t& v, int off) {
  fr_t r;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t lo = __shfl_down_sync(0xffffffffu, (uint32_t)v.l[i], off);
    uint32_t hi = __shfl_down_sync(0xffffffffu, (uint32_t)(v.l[i] >> 32), off);
    r.l[i] = ((uint64_t)hi << 32) | lo;
  }
  return r;
}
__global__ void __launch_bounds__(256) k_vec_sum(const fr_t* a, uint64_t n, fr_t* partial) {
  __shared__ fr_t sh[8];
  fr_t acc = zero<fr_t>();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    acc = add(acc, ld_fr(a + i));
  for (int off = 16; off >= 1; off >>= 1) acc = add(acc, shfl_down_fr(acc, off));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) acc = add(acc, sh[w]);
    st_fr(partial + blockIdx.x, acc);
  }
}

// out[rev(i)] = a[i] over log_n bits; in place (out == a) by swapping each pair once
// (reference caller: core/vecops.rs:392-535 -> icicle bit_reverse)
__global__ void __launch_bounds__(256) k_bit_reverse(const fr_t* a, uint32_t log_n, fr_t* out, bool in_place) {
  const uint64_t n = 1ull << log_n;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = log_n ? (__brevll(i) >> (64 - log_n)) : 0;
    if (!in_place) st_fr(out + r, ld_fr(a + i));
    else if (i < r) {
      fr_t x = ld_fr(a + i), y = ld_fr(a + r);
      st_fr(out + i, y);
      st_fr(out + r, x);
    }
  }
}

// standard <-> Montgomery form (reference: field_to_montgomery / field_from_montgomery, field.cuh:906-928)
template <bool TO_MONT>
__global__ void __launch_bounds__(256) k_mont_convert(const fr_t* a, uint64_t n, fr_t* out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    fr_t x = ld_fr(a + i);
    st_fr(out + i, TO_MONT ? to_mont(x) : from_mont(x));
  }
}

// out[i] = a[i]^-1 (0 -> 0): Montgomery's trick over M elements per thread, element j of thread t at index
// j*T + t (coalesced), one variable-time inversion per thread: 3 products per element + ~1/M of an inversion.
// (reference: batch_inv_cuda, vec_ops.cu:606-677, which falls back to one a^(r-2) per element above 2^18 elements)
constexpr int kInvM = 32;
__global__ void __launch_bounds__(128) k_vec_inv(const fr_t* a, uint64_t n, uint64_t T, fr_t* out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  fr_t pf[kInvM];
  fr_t acc = one<fr_t>();
#pragma unroll 1
  for (int j = 0; j < kInvM; j++) {
    const uint64_t idx = (uint64_t)j * T + t;
    pf[j] = acc;
    if (idx < n) {
      fr_t x = ld_fr(a + idx);
      if (!is_zero(x)) acc = mul(acc, x);
    }
  }
  fr_t inv_acc = inv_vartime(acc);
#pragma unroll 1
  for (int j = kInvM - 1; j >= 0; j--) {
    const uint64_t idx = (uint64_t)j * T + t;
    if (idx >= n) continue;
    fr_t x = ld_fr(a + idx);
    if (is_zero(x)) { st_fr(out + idx, x); continue; }
    st_fr(out + idx, mul(inv_acc, pf[j]));
    inv_acc = mul(inv_acc, x);
  }
}

enum UnaryOp { USUM = 0, UINV = 1, UBITREV = 2, UTOMONT = 3, UFROMMONT = 4 };

static unsigned stride_grid(uint64_t n) {
  uint64_t blocks = (n + 255) / 256;
  if (blocks > 148ull * 32) blocks = 148ull * 32;
  return (unsigned)(blocks ? blocks : 1);
}

static int unary_entry(int op, const b381_fr* a, uint64_t n, const b381_vecops_config* cfg, b381_fr* out) {
  if (!cfg) return B381_INVALID_POINTER;
  if (op == UBITREV && n && (n & (n - 1))) return B381_INVALID_ARGUMENT;     // power-of-two length only
  if (n == 0 && op != USUM) return B381_SUCCESS;
  if ((n && !a) || !out) return B381_INVALID_POINTER;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  {
    Scratch sc(st);
    const fr_t* da = nullptr;
    if ((e = stage_in(sc, (const fr_t*)a, n, cfg->is_a_on_device, &da)) != cudaSuccess) return map_cuda_error(e);
    const uint64_t n_out = op == USUM ? 1 : n;
    fr_t* dout = (fr_t*)out;
    if (!cfg->is_result_on_device && (e = sc.alloc(&dout, n_out)) != cudaSuccess) return map_cuda_error(e);
    if (op == USUM) {
      const unsigned g = stride_grid(n);
      fr_t* partial;
      if ((e = sc.alloc(&partial, (size_t)g)) != cudaSuccess) return map_cuda_error(e);
      k_vec_sum<<<g, 256, 0, st>>>(da, n, partial);
      k_vec_sum<<<1, 256, 0, st>>>(partial, g, dout);
    } else if (op == UINV) {
      const uint64_t T = (n + kInvM - 1) / kInvM;
      k_vec_inv<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(da, n, T, dout);
    } else if (op == UBITREV) {
      uint32_t log_n = 0;
      while ((1ull << log_n) < n) log_n++;
      k_bit_reverse<<<stride_grid(n), 256, 0, st>>>(da, log_n, dout, da == dout);
    } else if (op == UTOMONT) {
      k_mont_convert<true><<<stride_grid(n), 256, 0, st>>>(da, n, dout);
    } else {
      k_mont_convert<false><<<stride_grid(n), 256, 0, st>>>(da, n, dout);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->is_result_on_device) {
      e = cudaMemcpyAsync(out, dout, n_out * sizeof(fr_t), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return map_cuda_error(e);
    }
  }
  if (!cfg->is_async) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return map_cuda_error(e);
  }
  return B381_SUCCESS;
}

// The reference's flat `*_cuda` test entry points (vec_ops.cu:393-520): DEVICE vectors, output first, the
// scalar of scalar_*_vec_cuda read on the HOST; only `stream` of the config is honoured there.
template <int OP, bool A_SCALAR>
static int flat_entry(b381_fr* out, const b381_fr* a, const b381_fr* b, int n, const b381_vecops_config* cfg) {
  if (!cfg) return B381_INVALID_POINTER;
  if (n < 0) return B381_INVALID_ARGUMENT;
  b381_vecops_config c = *cfg;
  c.is_a_on_device = !A_SCALAR;
  c.is_b_on_device = true;
  c.is_result_on_device = true;
  c.is_async = false;
  return vecop_entry<OP, A_SCALAR>(a, b, (uint64_t)n, &c, out);
}

}  // namespace b381
using namespace b381;

extern "C" {
int b381_vector_add(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VADD, false>(a, b, n, c, o); }
int b381_vector_sub(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VSUB, false>(a, b, n, c, o); }
int b381_vector_mul(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VMUL, false>(a, b, n, c, o); }
int b381_scalar_mul_vec(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VMUL, true>(a, b, n, c, o); }
int b381_scalar_add_vec(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VADD, true>(a, b, n, c, o); }
int b381_scalar_mul_vec_batch(const b381_fr* a, const b381_fr* b, uint64_t n, int batch_size, bool columns_batch,
                              const b381_vecops_config* c, b381_fr* o) {
  return scalar_batch_entry<VMUL>(a, b, n, batch_size, columns_batch, c, o);
}
int b381_scalar_add_vec_batch(const b381_fr* a, const b381_fr* b, uint64_t n, int batch_size, bool columns_batch,
                              const b381_vecops_config* c, b381_fr* o) {
  return scalar_batch_entry<VADD>(a, b, n, batch_size, columns_batch, c, o);
}
int bls12_381_vector_add(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VADD, false>(a, b, n, c, o); }
int bls12_381_vector_sub(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VSUB, false>(a, b, n, c, o); }
int bls12_381_vector_mul(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VMUL, false>(a, b, n, c, o); }
int b381_vector_sum(const b381_fr* a, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return unary_entry(USUM, a, n, c, o); }
int b381_vector_inv(const b381_fr* a, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return unary_entry(UINV, a, n, c, o); }
int b381_bit_reverse(const b381_fr* a, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return unary_entry(UBITREV, a, n, c, o); }
int b381_montgomery_convert(const b381_fr* a, uint64_t n, int to_montgomery, const b381_vecops_config* c, b381_fr* o) {
  return unary_entry(to_montgomery ? UTOMONT : UFROMMONT, a, n, c, o);
}
int vec_add_cuda(b381_fr* o, const b381_fr* a, const b381_fr* b, int n, const b381_vecops_config* c) { return flat_entry<VADD, false>(o, a, b, n, c); }
int vec_sub_cuda(b381_fr* o, const b381_fr* a, const b381_fr* b, int n, const b381_vecops_config* c) { return flat_entry<VSUB, false>(o, a, b, n, c); }
int vec_mul_cuda(b381_fr* o, const b381_fr* a, const b381_fr* b, int n, const b381_vecops_config* c) { return flat_entry<VMUL, false>(o, a, b, n, c); }
int scalar_mul_vec_cuda(b381_fr* o, const b381_fr* s, const b381_fr* v, int n, const b381_vecops_config* c) { return flat_entry<VMUL, true>(o, s, v, n, c); }
int scalar_add_vec_cuda(b381_fr* o, const b381_fr* s, const b381_fr* v, int n, const b381_vecops_config* c) { return flat_entry<VADD, true>(o, s, v, n, c); }
int vec_sum_cuda(b381_fr* o, const b381_fr* a, int n, const b381_vecops_config* c) {
  if (!c) return B381_INVALID_POINTER;
  if (n < 0) return B381_INVALID_ARGUMENT;
  b381_vecops_config k = *c;
  k.is_a_on_device = true;            // the reference reads `is_result_on_device` for the one-element output (:506-510)
  k.is_async = false;
  return unary_entry(USUM, a, (uint64_t)n, &k, o);
}
b381_vecops_config b381_default_vecops_config(void) { b381_vecops_config c; memset(&c, 0, sizeof(c)); return c; }
}
