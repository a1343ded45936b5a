// Host entry points of msm_sort.cu: grouping of the (bucket, point) pairs, exclusive scan, task visiting order.
#pragma once
#include <functional>

#include "common.cuh"
#include "msm_core.cuh"

namespace b381 {

// out[i] = base + *base_dev + sum of in[0..i) for i < n (in == out allowed; base_dev may be null, and may alias out[0]);
// `copy`, when given, receives the same values (the scatter cursors); *total_dev, when given, receives *base_dev + the
// sum of all n (`base` not included).  `launches` counts the kernels launched.
cudaError_t exclusive_scan_u32(Scratch& sc, const uint32_t* in, uint32_t* out, size_t n, uint32_t* copy = nullptr,
                               int* launches = nullptr, uint32_t base = 0, const uint32_t* base_dev = nullptr,
                               uint32_t* total_dev = nullptr);

// Counting sort of the (run, point) pairs (run = chunk * nbuckets + bucket slot, msm_core.cuh): histogram, scan, scatter.
// hist: msm_runs(sh) + 1 words of scratch.  run_off[0 .. msm_runs(sh)] = run boundaries (last = n * W * batch),
// vals[n * W * batch] = entries (base index << 1 | sign) grouped by run, order inside a run unspecified.
// host_src != nullptr: the scalars are still in HOST memory; they are copied into d_scalars in pieces on a side stream
// and sorted as they land.
// after_piece(r0, r1, last): called in the STREAMED case (msm_sort_is_streamed: host scalars, chunk-major, one MSM,
// more than one piece) after the runs [r0, r1) of a piece have been scattered, in stream order -- run_off[r0 .. r1] is
// final then.  The last call has last = true and r1 = msm_runs(sh).
using msm_piece_fn = std::function<cudaError_t(size_t r0, size_t r1, bool last)>;
uint32_t msm_sort_piece(const msm_shape& sh);                       // scalars per copy / per streamed piece
bool msm_sort_is_streamed(const msm_shape& sh, const fr_t* host_src);
cudaError_t msm_sort_pairs(Scratch& sc, const fr_t* d_scalars, bool scalars_mont, const msm_shape& sh, const fr_t* host_src,
                           uint32_t* hist, uint32_t* run_off, uint32_t* vals, int* launches = nullptr,
                           const msm_piece_fn* after_piece = nullptr);

// order[0..ntasks) = task ids, longest task first (ties in unspecified order); K = task length bound
cudaError_t msm_task_order(Scratch& sc, size_t max_tasks, const uint32_t* ntasks_dev, const uint2* tasks, uint32_t K,
                           uint32_t* order, int* launches = nullptr);

}  // namespace b381
