// Window combine of the MSM: R = sum_w 2^(c*w) * S_w by Horner, c doublings per window -- ~255 doublings
// that nothing can shorten (the top window sum really has to be doubled 240 times), and a lone thread
// needs ~2400 cycles per Fq product (300 dependent carry-chain IMADs), 9 products per doubling: 3 ms.
// Here four lanes share each doubling: the XYZZ formulas (dbl-2008-s-1) have dependency depth 3
//   V = U^2, XX = X^2  |  W = U*V, S = X*V, ZZ' = V*ZZ, MM = M^2  |  M*(S-X3), W*Y, W*ZZZ
// so every lane multiplies one operand pair per level and the results travel by shuffle.
// Own translation unit (seconds to compile); also holds the finalize kernels.  Replaces final_accumulation_kernel
// (bls12-381/src/curve/msm_kernels.cu:529-596), whose thread W-1 does the same 240 doublings alone.
#include "common.cuh"
#include "msm_core.cuh"

namespace b381 {

template <class F>
__device__ __forceinline__ F shfl4(const F& v, int src) {
  F r;
  const uint32_t* in = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* out = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) out[i] = __shfl_sync(0xffffffffu, in[i], src, 4);
  return r;
}

template <class F>
__device__ __forceinline__ F sel4(int role, const F& a0, const F& a1, const F& a2, const F& a3) {
  return role == 0 ? a0 : role == 1 ? a1 : role == 2 ? a2 : a3;
}

// every lane holds the same point; lanes with (lane & 3) == role do the role's product of each level
template <class F>
__device__ __forceinline__ xyzz_t<F> xyzz_dbl_coop(const xyzz_t<F>& p, int role) {
  if (is_inf(p)) return p;
  const F u = dbl(p.y);
  F r = mul(sel4(role, u, p.x, u, u), sel4(role, u, p.x, u, u));
  const F v = shfl4(r, 0), xx = shfl4(r, 1);
  const F m = add(dbl(xx), xx);
  r = mul(sel4(role, u, p.x, v, m), sel4(role, v, v, p.zz, m));
  const F w = shfl4(r, 0), s = shfl4(r, 1), zz3 = shfl4(r, 2), mm = shfl4(r, 3);
  xyzz_t<F> o;
  o.x = sub(mm, dbl(s));
  r = mul(sel4(role, m, w, w, w), sel4(role, sub(s, o.x), p.y, p.zzz, p.y));
  o.y = sub(shfl4(r, 0), shfl4(r, 1));
  o.zz = zz3;
  o.zzz = shfl4(r, 2);
  return o;
}

template <class F>
__global__ void __launch_bounds__(32) k_msm_combine(const xyzz_t<F>* wsum, uint32_t stride, uint32_t W, uint32_t c,
                                                     xyzz_t<F>* out) {
  const int role = threadIdx.x & 3;
  xyzz_t<F> r = xyzz_identity<F>();
  for (int w = (int)W - 1; w >= 0; w--) {
    for (uint32_t k = 0; k < c; k++) r = xyzz_dbl_coop(r, role);
    xyzz_add(r, wsum[(size_t)w * stride]);
  }
  if (threadIdx.x == 0) *out = r;
}

// ---- finalize: bucket = sum of its task partials.  One thread per bucket for the usual 0..8 partials,
// one warp per bucket for the heavy ones (skewed scalar distributions; at most kMaxTasksPerBucket).
template <class F>
__global__ void __launch_bounds__(128) k_msm_finalize(uint32_t nbuckets, const uint32_t* task_start,
                                                      const uint32_t* counts, const xyzz_t<F>* partial,
                                                      xyzz_t<F>* buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  msm_finalize_body<F>(b, nbuckets, task_start, counts, partial, buckets, kFinalizeSerialMax);
}

template <class F>
__device__ __forceinline__ xyzz_t<F> shfl_down_point(const xyzz_t<F>& v, int off) {
  xyzz_t<F> r;
  const uint32_t* in = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* out = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(xyzz_t<F>) / 4); i++) out[i] = __shfl_down_sync(0xffffffffu, in[i], off);
  return r;
}

template <class F>
__global__ void __launch_bounds__(128) k_msm_finalize_heavy(uint32_t nbuckets, const uint32_t* task_start,
                                                            const uint32_t* counts, const xyzz_t<F>* partial,
                                                            xyzz_t<F>* buckets) {
  const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= nbuckets) return;
  const uint32_t nt = counts[b];
  if (nt <= kFinalizeSerialMax) return;
  const uint32_t t0 = task_start[b];
  xyzz_t<F> acc = xyzz_identity<F>();
  for (uint32_t t = lane; t < nt; t += 32) xyzz_add(acc, partial[t0 + t]);
  for (int off = 16; off >= 1; off >>= 1) {
    xyzz_t<F> o = shfl_down_point(acc, off);
    xyzz_add(acc, o);
  }
  if (lane == 0) buckets[b] = acc;
}

template <class F>
void launch_msm_finalize(uint32_t nbuckets, const uint32_t* task_start, const uint32_t* counts, const xyzz_t<F>* partial,
                         xyzz_t<F>* buckets, cudaStream_t st) {
  k_msm_finalize<F><<<grid_for(nbuckets, 128), 128, 0, st>>>(nbuckets, task_start, counts, partial, buckets);
  k_msm_finalize_heavy<F><<<grid_for((size_t)nbuckets * 32, 128), 128, 0, st>>>(nbuckets, task_start, counts, partial, buckets);
}
template void launch_msm_finalize<fq_t>(uint32_t, const uint32_t*, const uint32_t*, const xyzz_t<fq_t>*, xyzz_t<fq_t>*, cudaStream_t);
template void launch_msm_finalize<fq2_t>(uint32_t, const uint32_t*, const uint32_t*, const xyzz_t<fq2_t>*, xyzz_t<fq2_t>*, cudaStream_t);

template <class F>
void launch_msm_combine(const xyzz_t<F>* wsum, uint32_t stride, uint32_t W, uint32_t c, xyzz_t<F>* out, cudaStream_t st) {
  k_msm_combine<F><<<1, 32, 0, st>>>(wsum, stride, W, c, out);
}
template void launch_msm_combine<fq_t>(const xyzz_t<fq_t>*, uint32_t, uint32_t, uint32_t, xyzz_t<fq_t>*, cudaStream_t);
template void launch_msm_combine<fq2_t>(const xyzz_t<fq2_t>*, uint32_t, uint32_t, uint32_t, xyzz_t<fq2_t>*, cudaStream_t);

}  // namespace b381
