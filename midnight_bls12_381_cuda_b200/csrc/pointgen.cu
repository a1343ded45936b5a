// Arithmetic series of curve points on the device: out[i] = P0 + i*D (affine, Montgomery), G1 and G2.
// Used to lay down large synthetic base sets (SRS-like, all distinct, known discrete logs when
// P0 = a*G and D = b*G) directly in HBM -- e.g. bench.py's 2^24 bases -- and as the first consumer of
// the batched XYZZ->affine conversion (Montgomery's trick: one inversion per 16 points), which is the
// "next" row 3 of SURVEY.md 8(f); the reference converts with one inversion per point
// (bls12-381/src/curve/point_ops.cu:61-101, :354-547).
#include "common.cuh"
#include "msm_core.cuh"

namespace b381 {

constexpr int kSeriesChunk = 16;

template <class F>
__global__ void __launch_bounds__(64) k_point_series(affine_t<F> p0, affine_t<F> d, uint64_t n, affine_t<F>* out) {
  uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i0 = chunk * kSeriesChunk;
  if (i0 >= n) return;
  // start = P0 + i0 * D  (double-and-add over the bits of i0)
  xyzz_t<F> dx = to_xyzz(d);
  xyzz_t<F> cur = xyzz_identity<F>();
  for (int b = 63; b >= 0; b--) {
    cur = xyzz_dbl(cur);
    if ((i0 >> b) & 1) xyzz_add(cur, dx);
  }
  if (!is_inf(p0)) xyzz_madd(cur, p0);
  xyzz_t<F> pts[kSeriesChunk];
  F pre[kSeriesChunk];
  F acc = one<F>();
  int m = 0;
  for (; m < kSeriesChunk && i0 + m < n; m++) {
    pts[m] = cur;
    pre[m] = acc;
    if (!is_inf(cur)) acc = mul(acc, cur.zzz);
    if (!is_inf(d)) xyzz_madd(cur, d);
  }
  F inv_acc = inv(acc);
  for (int j = m - 1; j >= 0; j--) {
    affine_t<F> a;
    if (is_inf(pts[j])) {
      a = affine_t<F>{zero<F>(), zero<F>()};
    } else {
      F iz3 = mul(inv_acc, pre[j]);          // 1/ZZZ_j
      inv_acc = mul(inv_acc, pts[j].zzz);
      F t = mul(iz3, pts[j].zz);             // ZZ/ZZZ
      F iz2 = sqr(t);                        // 1/ZZ
      a = affine_t<F>{mul(pts[j].x, iz2), mul(pts[j].y, iz3)};
    }
    out[i0 + j] = a;
  }
}

template <class F>
static int series_entry(const affine_t<F>* p0, const affine_t<F>* d, uint64_t n, affine_t<F>* out, void* stream) {
  if (!p0 || !d || (!out && n)) return B381_INVALID_POINTER;
  if (n == 0) return B381_SUCCESS;
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t chunks = (n + kSeriesChunk - 1) / kSeriesChunk;
  k_point_series<F><<<grid_for(chunks, 64), 64, 0, st>>>(*p0, *d, n, out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return map_cuda_error(e);
}

}  // namespace b381
using namespace b381;

extern "C" {
int b381_g1_point_series(const b381_g1_affine* p0, const b381_g1_affine* d, uint64_t n, b381_g1_affine* out_device, void* stream) {
  return series_entry<fq_t>((const g1_affine*)p0, (const g1_affine*)d, n, (g1_affine*)out_device, stream);
}
int b381_g2_point_series(const b381_g2_affine* p0, const b381_g2_affine* d, uint64_t n, b381_g2_affine* out_device, void* stream) {
  return series_entry<fq2_t>((const g2_affine*)p0, (const g2_affine*)d, n, (g2_affine*)out_device, stream);
}
}
