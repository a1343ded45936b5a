// Host entry points of msm_sort.cu: grouping of the (bucket, point) pairs, exclusive scan, task visiting order.
#pragma once
#include "common.cuh"
#include "msm_core.cuh"

namespace b381 {

// out[i] = base + sum of in[0..i) for i < n (in == out allowed); `copy`, when given, receives the same values (the
// scatter cursors).  `launches` counts the kernels launched.
cudaError_t exclusive_scan_u32(Scratch& sc, const uint32_t* in, uint32_t* out, size_t n, uint32_t* copy = nullptr,
                               int* launches = nullptr, uint32_t base = 0);

// Counting sort of the (run, point) pairs (run = chunk * nbuckets + bucket slot, msm_core.cuh): histogram, scan, scatter.
// hist: msm_runs(sh) + 1 words of scratch.  run_off[0 .. msm_runs(sh)] = run boundaries (last = n * W * batch),
// vals[n * W * batch] = entries (base index << 1 | sign) grouped by run, order inside a run unspecified.
// host_src != nullptr: the scalars are still in HOST memory; they are copied into d_scalars in pieces on a side stream
// and sorted as they land.
cudaError_t msm_sort_pairs(Scratch& sc, const fr_t* d_scalars, bool scalars_mont, const msm_shape& sh, const fr_t* host_src,
                           uint32_t* hist, uint32_t* run_off, uint32_t* vals, int* launches = nullptr);

// order[0..ntasks) = task ids, longest task first (ties in unspecified order); K = task length bound
cudaError_t msm_task_order(Scratch& sc, size_t max_tasks, const uint32_t* ntasks_dev, const uint2* tasks, uint32_t K,
                           uint32_t* order, int* launches = nullptr);

}  // namespace b381
