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

// ---- cooperation policies -------------------------------------------------------------------------------------
// A group of lanes holds the same point(s); every dependency level of a group-law formula is ONE field product per
// lane, and the results travel by shuffle.
//   coop4<F>   4 lanes per group: lane `role` = lane & 3 multiplies the role's operand pair of the level.
//   coop16     16 lanes per group, Fq2 only: role = (lane >> 2) & 3, and the role's Fq2 product is itself spread over
//              its four sub-lanes (Karatsuba: a0 b0 | a1 b1 | (a0+a1)(b0+b1) | idle), so a level costs one Fq product
//              instead of three.  For the ONE dependency chain of the window combine (255 doublings): G2 2^20
//              2.95 -> 1.52 ms.
template <class F>
__device__ __forceinline__ F shfl_words(const F& v, int src, int width) {
  F r;
  const uint32_t* in = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* out = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) out[i] = __shfl_sync(0xffffffffu, in[i], src, width);
  return r;
}

template <class F>
__device__ __forceinline__ F sel4(int role, const F& a0, const F& a1, const F& a2, const F& a3) {
  return role == 0 ? a0 : role == 1 ? a1 : role == 2 ? a2 : a3;
}

template <class F>
struct coop4 {
  static constexpr int G = 4;
  int role;
  __device__ __forceinline__ coop4() : role(threadIdx.x & 3) {}
  __device__ __forceinline__ F cmul(const F& a, const F& b) const { return mul(a, b); }
  __device__ __forceinline__ F from(const F& v, int r) const { return shfl_words(v, r, 4); }
};

struct coop16 {
  static constexpr int G = 16;
  int role, sub;
  __device__ __forceinline__ coop16() : role((threadIdx.x >> 2) & 3), sub(threadIdx.x & 3) {}
  __device__ __forceinline__ fq2_t cmul(const fq2_t& a, const fq2_t& b) const {
    const fq_t x = sub == 0 ? a.c0 : sub == 1 ? a.c1 : add(a.c0, a.c1);
    const fq_t y = sub == 0 ? b.c0 : sub == 1 ? b.c1 : add(b.c0, b.c1);
    const fq_t r = mul(x, y);
    const fq_t v0 = shfl_words(r, 0, 4), v1 = shfl_words(r, 1, 4), s = shfl_words(r, 2, 4);
    return fq2_t{b381::sub(v0, v1), b381::sub(b381::sub(s, v0), v1)};
  }
  // every sub-lane of role r holds r's product: take it from the sub-lane that sits where this lane does
  __device__ __forceinline__ fq2_t from(const fq2_t& v, int r) const { return shfl_words(v, r * 4 + sub, 16); }
};
template <class F> struct coop_for { using wide = coop4<F>; };
template <> struct coop_for<fq2_t> { using wide = coop16; };

// every lane of the group holds the same point; the lanes of role r do the role's product of each level
template <class F, class C>
__device__ __forceinline__ xyzz_t<F> xyzz_dbl_coop_always(const xyzz_t<F>& p, const C& co) {
  const int role = co.role;
  const F u = dbl(p.y);
  F r = co.cmul(sel4(role, u, p.x, u, u), sel4(role, u, p.x, u, u));
  const F v = co.from(r, 0), xx = co.from(r, 1);
  const F m = add(dbl(xx), xx);
  r = co.cmul(sel4(role, u, p.x, v, m), sel4(role, v, v, p.zz, m));
  const F w = co.from(r, 0), s = co.from(r, 1), zz3 = co.from(r, 2), mm = co.from(r, 3);
  xyzz_t<F> o;
  o.x = sub(mm, dbl(s));
  r = co.cmul(sel4(role, m, w, w, w), sel4(role, sub(s, o.x), p.y, p.zzz, p.y));
  o.y = sub(co.from(r, 0), co.from(r, 1));
  o.zz = zz3;
  o.zzz = co.from(r, 2);
  return o;
}
// callers whose whole warp holds the same point (the combine kernel) may skip the identity
template <class F, class C>
__device__ __forceinline__ xyzz_t<F> xyzz_dbl_coop(const xyzz_t<F>& p, const C& co) {
  if (is_inf(p)) return p;
  return xyzz_dbl_coop_always(p, co);
}

// ---- bucket reduction: window_sum[w] = sum_j (j+1) * bucket[w][j]  (msm_core.cuh: msm_segment_body /
// msm_tree_body are the serial statements of the same sums, used by the CPU-only harness).  Here a GROUP of lanes
// owns a segment: the general XYZZ addition (add-2008-s, 12M+2S) has dependency depth 4
//   U1,U2,S1,S2 | PP,RR,ZZ1*ZZ2,ZZZ1*ZZZ2 | PPP,Q,ZZ3 | R*(Q-X3),S1*PPP,ZZZ3
// so a lane multiplies one operand pair per level and results travel by shuffles: the 2L dependent
// additions of a segment cost 4 product latencies each instead of 14.  All lanes of a group hold the same
// points, so the special cases (identity operands, P = +-Q) are resolved by selects after the common path
// and every lane always executes the same shuffles.
template <class F, class C>
__device__ __forceinline__ void xyzz_add_coop(xyzz_t<F>& acc, const xyzz_t<F>& q, const C& co) {
  const int role = co.role;
  const bool q_inf = is_inf(q), a_inf = is_inf(acc);
  F r = co.cmul(sel4(role, acc.x, q.x, acc.y, q.y), sel4(role, q.zz, acc.zz, q.zzz, acc.zzz));
  const F u1 = co.from(r, 0), u2 = co.from(r, 1), s1 = co.from(r, 2), s2 = co.from(r, 3);
  const F p = sub(u2, u1), rr0 = sub(s2, s1);
  r = co.cmul(sel4(role, p, rr0, acc.zz, acc.zzz), sel4(role, p, rr0, q.zz, q.zzz));
  const F pp = co.from(r, 0), rr = co.from(r, 1), z12 = co.from(r, 2), z123 = co.from(r, 3);
  r = co.cmul(sel4(role, p, u1, z12, z12), pp);
  const F ppp = co.from(r, 0), q1 = co.from(r, 1), zz3 = co.from(r, 2);
  xyzz_t<F> o;
  o.x = sub(sub(rr, ppp), dbl(q1));
  r = co.cmul(sel4(role, rr0, s1, z123, z123), sel4(role, sub(q1, o.x), ppp, ppp, ppp));
  o.y = sub(co.from(r, 0), co.from(r, 1));
  o.zz = zz3;
  o.zzz = co.from(r, 2);
  if (q_inf) return;
  if (a_inf) { acc = q; return; }
  if (is_zero(p)) {                       // same x: doubling or cancellation (rare), serial formulas
    if (is_zero(rr0)) acc = xyzz_dbl(acc);
    else acc = xyzz_identity<F>();
    return;
  }
  acc = o;
}

// k * p for k < 2^kbits, the products shared by the lanes of the group; all lanes of a group see the same k
template <class F, class C>
__device__ __forceinline__ xyzz_t<F> xyzz_mul_small_coop(const xyzz_t<F>& p, uint32_t k, int kbits, const C& co) {
  xyzz_t<F> r = xyzz_identity<F>();
#pragma unroll 1
  for (int i = kbits - 1; i >= 0; i--) {
    // every group of the warp runs all kbits rounds so that the shuffles stay warp-uniform
    xyzz_t<F> d = r;
    {
      // xyzz_dbl_coop returns early on the identity; run its shuffles unconditionally instead
      const bool inf = is_inf(r);
      xyzz_t<F> t = r;
      if (inf) t = p;                    // any finite point keeps the arithmetic well defined
      t = xyzz_dbl_coop_always(t, co);
      if (!inf) d = t;
    }
    r = d;
    xyzz_t<F> s = r;
    xyzz_add_coop(s, p, co);
    if ((k >> i) & 1) r = s;
  }
  return r;
}

template <class F, class C>
__global__ void __launch_bounds__(128) k_msm_segment_coop(uint32_t W, uint32_t B, uint32_t L, const xyzz_t<F>* buckets,
                                                          xyzz_t<F>* seg_out) {
  const C co;
  const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / C::G;
  const bool lead = (threadIdx.x & (C::G - 1)) == 0;
  const uint32_t segs = B / L;
  const bool live = gid < W * segs;       // dead groups still run the shuffles (on segment 0)
  const uint32_t g = live ? gid : 0;
  const uint32_t w = g / segs, s = g % segs;
  const xyzz_t<F>* bk = buckets + (size_t)w * (B + 1) + (size_t)s * L;     // B + 1 slots per set (trash last)
  xyzz_t<F> run = xyzz_identity<F>();
  xyzz_t<F> tri = xyzz_identity<F>();
#pragma unroll 1
  for (int j = (int)L - 1; j >= 0; j--) {
    xyzz_add_coop(run, bk[j], co);
    xyzz_add_coop(tri, run, co);
  }
  xyzz_t<F> sh = xyzz_mul_small_coop(run, s * L, 32 - __clz((B - 1) | 1), co);
  xyzz_add_coop(tri, sh, co);
  if (live && lead) seg_out[g] = tri;
}

// in-place halving over groups: a[g*stride + t] += a[g*stride + t + half]
template <class F, class C>
__global__ void __launch_bounds__(128) k_msm_tree_coop(uint32_t groups, uint32_t stride, uint32_t half, xyzz_t<F>* a) {
  const C co;
  const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / C::G;
  const bool lead = (threadIdx.x & (C::G - 1)) == 0;
  const bool live = gid < groups * half;
  const uint32_t i = live ? gid : 0;
  const uint32_t g = i / half, t = i % half;
  xyzz_t<F> x = a[(size_t)g * stride + t];
  xyzz_add_coop(x, a[(size_t)g * stride + t + half], co);
  if (live && lead) a[(size_t)g * stride + t] = x;
}

// seg (W * B/L entries) is scratch; on return seg[w * (B/L)] = window sum w
template <class F>
void launch_msm_bucket_reduce(uint32_t W, uint32_t B, uint32_t L, const xyzz_t<F>* buckets, xyzz_t<F>* seg, cudaStream_t st) {
  // four lanes per segment for G2 as well: with thousands of segments in flight this stage is bound by the number of
  // products executed, not by their latency, and coop16 executes more of them (G2 2^20: 4.2 ms -> 11.4 ms, measured)
  using C = coop4<F>;
  const uint32_t segs = B / L;
  k_msm_segment_coop<F, C><<<grid_for((size_t)W * segs * C::G, 128), 128, 0, st>>>(W, B, L, buckets, seg);
  for (uint32_t half = segs / 2; half >= 1; half >>= 1)
    k_msm_tree_coop<F, C><<<grid_for((size_t)W * half * C::G, 128), 128, 0, st>>>(W, segs, half, seg);
}
template void launch_msm_bucket_reduce<fq_t>(uint32_t, uint32_t, uint32_t, const xyzz_t<fq_t>*, xyzz_t<fq_t>*, cudaStream_t);
template void launch_msm_bucket_reduce<fq2_t>(uint32_t, uint32_t, uint32_t, const xyzz_t<fq2_t>*, xyzz_t<fq2_t>*, cudaStream_t);

template <class F, class C>
__global__ void __launch_bounds__(32) k_msm_combine(const xyzz_t<F>* wsum, uint32_t stride, uint32_t W, uint32_t c,
                                                     xyzz_t<F>* out) {
  const C co;
  wsum += (size_t)blockIdx.x * W * stride;            // one CTA per MSM of the batch
  xyzz_t<F> r = xyzz_identity<F>();
  for (int w = (int)W - 1; w >= 0; w--) {
    for (uint32_t k = 0; k < c; k++) r = xyzz_dbl_coop(r, co);
    xyzz_add_coop(r, wsum[(size_t)w * stride], co);
  }
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

// ---- finalize: bucket = sum of its task partials.  One thread per bucket for the usual 0..8 partials; buckets with
// more (skewed scalar distributions; at most kMaxTasksPerBucket) are appended to a list and summed by one warp each.
// The list keeps the second kernel's grid small: as a warp-per-bucket launch over ALL buckets it spent 0.28 ms of a
// 2^24-point MSM scheduling 131 k CTAs that found nothing to do (profiles/r01d_launches_summary.txt).
template <class F>
__global__ void __launch_bounds__(128) k_msm_finalize(uint32_t nbuckets, const uint32_t* task_start,
                                                      const uint32_t* counts, const xyzz_t<F>* partial,
                                                      xyzz_t<F>* buckets, uint32_t* heavy_list, uint32_t* heavy_count) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbuckets && counts[b] > kFinalizeSerialMax) heavy_list[atomicAdd(heavy_count, 1u)] = b;
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

constexpr unsigned kHeavyGrid = 148 * 4;       // warps loop over the list

template <class F>
__global__ void __launch_bounds__(128) k_msm_finalize_heavy(const uint32_t* heavy_list, const uint32_t* heavy_count,
                                                            const uint32_t* task_start, const uint32_t* counts,
                                                            const xyzz_t<F>* partial, xyzz_t<F>* buckets) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5, total = *heavy_count;
  for (uint32_t i = warp; i < total; i += nwarps) {
    const uint32_t b = heavy_list[i];
    const uint32_t nt = counts[b], t0 = task_start[b];
    xyzz_t<F> acc = xyzz_identity<F>();
    for (uint32_t t = lane; t < nt; t += 32) xyzz_add(acc, partial[t0 + t]);
    for (int off = 16; off >= 1; off >>= 1) {
      xyzz_t<F> o = shfl_down_point(acc, off);
      xyzz_add(acc, o);
    }
    if (lane == 0) buckets[b] = acc;
  }
}

// heavy_list: nbuckets entries of scratch; heavy_count: one zeroed word
template <class F>
void launch_msm_finalize(uint32_t nbuckets, const uint32_t* task_start, const uint32_t* counts, const xyzz_t<F>* partial,
                         xyzz_t<F>* buckets, uint32_t* heavy_list, uint32_t* heavy_count, cudaStream_t st) {
  cudaMemsetAsync(heavy_count, 0, sizeof(uint32_t), st);
  k_msm_finalize<F><<<grid_for(nbuckets, 128), 128, 0, st>>>(nbuckets, task_start, counts, partial, buckets, heavy_list, heavy_count);
  k_msm_finalize_heavy<F><<<kHeavyGrid, 128, 0, st>>>(heavy_list, heavy_count, task_start, counts, partial, buckets);
}
template void launch_msm_finalize<fq_t>(uint32_t, const uint32_t*, const uint32_t*, const xyzz_t<fq_t>*, xyzz_t<fq_t>*, uint32_t*, uint32_t*, cudaStream_t);
template void launch_msm_finalize<fq2_t>(uint32_t, const uint32_t*, const uint32_t*, const xyzz_t<fq2_t>*, xyzz_t<fq2_t>*, uint32_t*, uint32_t*, cudaStream_t);

template <class F>
void launch_msm_combine(const xyzz_t<F>* wsum, uint32_t stride, uint32_t W, uint32_t c, xyzz_t<F>* out, uint32_t batch,
                        cudaStream_t st) {
  k_msm_combine<F, typename coop_for<F>::wide><<<batch, 32, 0, st>>>(wsum, stride, W, c, out);
}
template void launch_msm_combine<fq_t>(const xyzz_t<fq_t>*, uint32_t, uint32_t, uint32_t, xyzz_t<fq_t>*, uint32_t, cudaStream_t);
template void launch_msm_combine<fq2_t>(const xyzz_t<fq2_t>*, uint32_t, uint32_t, uint32_t, xyzz_t<fq2_t>*, uint32_t, cudaStream_t);

}  // namespace b381
