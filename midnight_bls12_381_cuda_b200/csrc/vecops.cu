// Fr element-wise vector operations (Montgomery domain), 128-bit coalesced accesses.
// Replaces vec_{add,sub,mul}_kernel / scalar_vec_{mul,add}_kernel (bls12-381/src/field/vec_ops.cu:63-118,
// :335-345) and the run_vec_op staging helper (bls12-381/src/backend/icicle_field_api.cu:133-334).
// HBM-bound: 96 B/element (two reads + one write), 64 B/element for scalar (op) vector.
#include <cstring>

#include "common.cuh"
#include "field.cuh"

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

}  // namespace b381
using namespace b381;

extern "C" {
int b381_vector_add(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VADD, false>(a, b, n, c, o); }
int b381_vector_sub(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VSUB, false>(a, b, n, c, o); }
int b381_vector_mul(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VMUL, false>(a, b, n, c, o); }
int b381_scalar_mul_vec(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VMUL, true>(a, b, n, c, o); }
int b381_scalar_add_vec(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VADD, true>(a, b, n, c, o); }
int bls12_381_vector_add(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VADD, false>(a, b, n, c, o); }
int bls12_381_vector_sub(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VSUB, false>(a, b, n, c, o); }
int bls12_381_vector_mul(const b381_fr* a, const b381_fr* b, uint64_t n, const b381_vecops_config* c, b381_fr* o) { return vecop_entry<VMUL, false>(a, b, n, c, o); }
b381_vecops_config b381_default_vecops_config(void) { b381_vecops_config c; memset(&c, 0, sizeof(c)); return c; }
}
