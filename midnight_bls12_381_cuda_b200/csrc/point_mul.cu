// Batch scalar multiplication on G1 (GLV and plain) and subgroup-membership checks for G1 and G2
// (bodies in glv.cuh).
//   bls12_381_g1_scalar_mul_glv / bls12_381_g1_scalar_mul    names, argument order, residency flags and error codes of
//                                                            bls12-381/src/curve/point_ops.cu:1019-1268
//   b381_g1_is_in_subgroup / b381_g2_is_in_subgroup          the checks include/point.cuh:419-448 leaves as TODO
// Contract kept from the reference: bases Montgomery affine, scalars CANONICAL integers (the kernels read the limbs as
// they are, point_ops.cu:364, :491), output Jacobian Montgomery.  The Jacobian representative is not unique; here it is
// always the normalised one (x, y, 1), identity (0, R, 0), so outputs can be compared byte for byte.
// Not kept: the reference's constant-time table scan -- prover inputs are public, lookups here are indexed.
#include "common.cuh"
#include "glv.cuh"

namespace b381 {

template <bool GLV>
__global__ void __launch_bounds__(128) k_g1_scalar_mul(const g1_affine* bases, const fr_t* scalars, uint32_t n, g1_jac* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const g1_affine p = bases[i];
  const fr_t k = scalars[i];
  const g1_xyzz r = GLV ? g1_mul_glv(p, k) : g1_mul_window(p, k);
  g1_jac o;
  if (is_inf(r)) { o.x = zero<fq_t>(); o.y = one<fq_t>(); o.z = zero<fq_t>(); }
  else { const g1_affine a = xyzz_to_affine(r); o.x = a.x; o.y = a.y; o.z = one<fq_t>(); }
  out[i] = o;
}

__global__ void __launch_bounds__(128) k_g1_in_subgroup(const g1_affine* in, uint32_t n, uint8_t* flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = g1_in_subgroup(in[i]) ? 1 : 0;
}
__global__ void __launch_bounds__(128) k_g2_in_subgroup(const g2_affine* in, uint32_t n, uint8_t* flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = g2_in_subgroup(in[i]) ? 1 : 0;
}

constexpr int kMaxPointBatch = 1 << 26;

static int scalar_mul_entry(bool glv, const g1_affine* bases, const fr_t* scalars, int size, const b381_vecops_config* cfg,
                            g1_jac* out) {
  if (!bases || !scalars || !out || !cfg) return B381_INVALID_ARGUMENT;      // the reference's codes (point_ops.cu:1027-1035)
  if (size <= 0 || size > kMaxPointBatch) return B381_INVALID_ARGUMENT;
  const uint32_t n = (uint32_t)size;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  {
    Scratch sc(st);
    const g1_affine* d_b;
    const fr_t* d_s;
    if ((e = stage_in(sc, bases, n, cfg->is_a_on_device, &d_b)) != cudaSuccess) return map_cuda_error(e);
    if ((e = stage_in(sc, scalars, n, cfg->is_b_on_device, &d_s)) != cudaSuccess) return map_cuda_error(e);
    g1_jac* d_o = out;
    if (!cfg->is_result_on_device && (e = sc.alloc(&d_o, n)) != cudaSuccess) return map_cuda_error(e);
    if (glv) k_g1_scalar_mul<true><<<grid_for(n, 128), 128, 0, st>>>(d_b, d_s, n, d_o);
    else k_g1_scalar_mul<false><<<grid_for(n, 128), 128, 0, st>>>(d_b, d_s, n, d_o);
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->is_result_on_device &&
        (e = cudaMemcpyAsync(out, d_o, sizeof(g1_jac) * n, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
      return map_cuda_error(e);
  }
  if (!cfg->is_async && (e = cudaStreamSynchronize(st)) != cudaSuccess) return map_cuda_error(e);
  return B381_SUCCESS;
}

static void launch_subgroup(const g1_affine* in, uint32_t n, uint8_t* f, cudaStream_t st) {
  k_g1_in_subgroup<<<grid_for(n, 128), 128, 0, st>>>(in, n, f);
}
static void launch_subgroup(const g2_affine* in, uint32_t n, uint8_t* f, cudaStream_t st) {
  k_g2_in_subgroup<<<grid_for(n, 128), 128, 0, st>>>(in, n, f);
}

template <class A>
static int subgroup_entry(const A* in, int size, const b381_vecops_config* cfg, uint8_t* flags) {
  if (!in || !flags || !cfg) return B381_INVALID_ARGUMENT;
  if (size <= 0 || size > kMaxPointBatch) return B381_INVALID_ARGUMENT;
  const uint32_t n = (uint32_t)size;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  {
    Scratch sc(st);
    const A* d_in;
    if ((e = stage_in(sc, in, n, cfg->is_a_on_device, &d_in)) != cudaSuccess) return map_cuda_error(e);
    uint8_t* d_f = flags;
    if (!cfg->is_result_on_device && (e = sc.alloc(&d_f, n)) != cudaSuccess) return map_cuda_error(e);
    launch_subgroup(d_in, n, d_f, st);
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->is_result_on_device && (e = cudaMemcpyAsync(flags, d_f, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
      return map_cuda_error(e);
  }
  if (!cfg->is_async && (e = cudaStreamSynchronize(st)) != cudaSuccess) return map_cuda_error(e);
  return B381_SUCCESS;
}

}  // namespace b381
using namespace b381;

extern "C" {
int bls12_381_g1_scalar_mul_glv(const b381_g1_affine* bases, const b381_fr* scalars, int size, const b381_vecops_config* cfg,
                                b381_g1_projective* out) {
  return scalar_mul_entry(true, (const g1_affine*)bases, (const fr_t*)scalars, size, cfg, (g1_jac*)out);
}
int bls12_381_g1_scalar_mul(const b381_g1_affine* bases, const b381_fr* scalars, int size, const b381_vecops_config* cfg,
                            b381_g1_projective* out) {
  return scalar_mul_entry(false, (const g1_affine*)bases, (const fr_t*)scalars, size, cfg, (g1_jac*)out);
}
int b381_g1_is_in_subgroup(const b381_g1_affine* in, int size, const b381_vecops_config* cfg, uint8_t* flags) {
  return subgroup_entry((const g1_affine*)in, size, cfg, flags);
}
int b381_g2_is_in_subgroup(const b381_g2_affine* in, int size, const b381_vecops_config* cfg, uint8_t* flags) {
  return subgroup_entry((const g2_affine*)in, size, cfg, flags);
}
}
