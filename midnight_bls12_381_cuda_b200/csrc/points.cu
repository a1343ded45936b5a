// Point-format conversions and ingest validation for G1 and G2:
//   affine -> Jacobian            (bls12-381/src/curve/point_ops.cu:61-73,  exported :759-842)
//   Jacobian -> affine, batched   (point_ops.cu:78-101, exported :844-1000; the reference inverts every Z on its
//                                  own with a^(p-2); here Montgomery's trick shares one variable-time inversion
//                                  between kConvM points: 6 products per point instead of ~570)
//   on-curve check                (g1_is_on_curve / g2_is_on_curve, include/point.cuh:339-387)
// Wire formats are the reference's: affine (x, y) Montgomery with infinity = (0, 0) (point.cuh:295-302);
// Jacobian (X, Y, Z) Montgomery, x = X/Z^2, y = Y/Z^3, infinity = (0, R, 0) (point.cuh:469-486).
#include "common.cuh"
#include "curve.cuh"

namespace b381 {

constexpr int kConvM = 16;

template <class F>
__global__ void __launch_bounds__(256) k_affine_to_jac(const affine_t<F>* in, uint64_t n, jacobian_t<F>* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  affine_t<F> p = in[i];
  jacobian_t<F> o;
  if (is_inf(p)) { o.x = zero<F>(); o.y = one<F>(); o.z = zero<F>(); }
  else { o.x = p.x; o.y = p.y; o.z = one<F>(); }
  out[i] = o;
}

// thread t converts points t, t + T, t + 2T, ... (coalesced across the warp)
template <class F>
__global__ void __launch_bounds__(128) k_jac_to_affine(const jacobian_t<F>* in, uint64_t n, uint64_t T, affine_t<F>* out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  F pre[kConvM];
  F acc = one<F>();
#pragma unroll 1
  for (int j = 0; j < kConvM; j++) {
    const uint64_t idx = (uint64_t)j * T + t;
    pre[j] = acc;
    if (idx < n) {
      F z = in[idx].z;
      if (!is_zero(z)) acc = mul(acc, z);
    }
  }
  F inv_acc = inv_vartime(acc);
#pragma unroll 1
  for (int j = kConvM - 1; j >= 0; j--) {
    const uint64_t idx = (uint64_t)j * T + t;
    if (idx >= n) continue;
    jacobian_t<F> p = in[idx];
    if (is_zero(p.z)) { out[idx] = affine_t<F>{zero<F>(), zero<F>()}; continue; }
    F zi = mul(inv_acc, pre[j]);
    inv_acc = mul(inv_acc, p.z);
    F zi2 = sqr(zi);
    out[idx] = affine_t<F>{mul(p.x, zi2), mul(p.y, mul(zi2, zi))};
  }
}

// flags[i] = 1 when y^2 = x^3 + b or the point is infinity, else 0
template <class F>
__global__ void __launch_bounds__(256) k_on_curve(const affine_t<F>* in, uint64_t n, uint8_t* flags) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  affine_t<F> p = in[i];
  flags[i] = on_curve(p) ? 1 : 0;
}

enum PointOp { P_TO_JAC = 0, P_TO_AFFINE = 1, P_ON_CURVE = 2 };

template <class F>
static int point_entry(int op, const void* in, int size, const b381_vecops_config* cfg, void* out) {
  if (!in || !out || !cfg) return B381_INVALID_ARGUMENT;       // the reference's codes (point_ops.cu:766-775)
  if (size <= 0 || size > (1 << 26)) return B381_INVALID_ARGUMENT;
  const uint64_t n = (uint64_t)size;
  const size_t in_sz = op == P_TO_AFFINE ? sizeof(jacobian_t<F>) : sizeof(affine_t<F>);
  const size_t out_sz = op == P_TO_JAC ? sizeof(jacobian_t<F>) : op == P_TO_AFFINE ? sizeof(affine_t<F>) : 1;
  cudaStream_t st = (cudaStream_t)cfg->stream;
  cudaError_t e;
  {
    Scratch sc(st);
    const uint8_t* d_in;
    if ((e = stage_in(sc, (const uint8_t*)in, n * in_sz, cfg->is_a_on_device, &d_in)) != cudaSuccess) return map_cuda_error(e);
    uint8_t* d_out = (uint8_t*)out;
    if (!cfg->is_result_on_device && (e = sc.alloc(&d_out, n * out_sz)) != cudaSuccess) return map_cuda_error(e);
    if (op == P_TO_JAC) {
      k_affine_to_jac<F><<<grid_for(n, 256), 256, 0, st>>>((const affine_t<F>*)d_in, n, (jacobian_t<F>*)d_out);
    } else if (op == P_TO_AFFINE) {
      const uint64_t T = (n + kConvM - 1) / kConvM;
      k_jac_to_affine<F><<<grid_for(T, 128), 128, 0, st>>>((const jacobian_t<F>*)d_in, n, T, (affine_t<F>*)d_out);
    } else {
      k_on_curve<F><<<grid_for(n, 256), 256, 0, st>>>((const affine_t<F>*)d_in, n, d_out);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return map_cuda_error(e);
    if (!cfg->is_result_on_device) {
      e = cudaMemcpyAsync(out, d_out, n * out_sz, cudaMemcpyDeviceToHost, st);
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
int bls12_381_g1_affine_to_projective(const b381_g1_affine* in, int size, const b381_vecops_config* cfg, b381_g1_projective* out) {
  return point_entry<fq_t>(P_TO_JAC, in, size, cfg, out);
}
int bls12_381_g1_projective_to_affine(const b381_g1_projective* in, int size, const b381_vecops_config* cfg, b381_g1_affine* out) {
  return point_entry<fq_t>(P_TO_AFFINE, in, size, cfg, out);
}
int bls12_381_g2_affine_to_projective(const b381_g2_affine* in, int size, const b381_vecops_config* cfg, b381_g2_projective* out) {
  return point_entry<fq2_t>(P_TO_JAC, in, size, cfg, out);
}
int bls12_381_g2_projective_to_affine(const b381_g2_projective* in, int size, const b381_vecops_config* cfg, b381_g2_affine* out) {
  return point_entry<fq2_t>(P_TO_AFFINE, in, size, cfg, out);
}
int b381_g1_is_on_curve(const b381_g1_affine* in, int size, const b381_vecops_config* cfg, uint8_t* flags) {
  return point_entry<fq_t>(P_ON_CURVE, in, size, cfg, flags);
}
int b381_g2_is_on_curve(const b381_g2_affine* in, int size, const b381_vecops_config* cfg, uint8_t* flags) {
  return point_entry<fq2_t>(P_ON_CURVE, in, size, cfg, flags);
}
}
