// Host entry points of msm_sort.cu: grouping of the (bucket, point) pairs, exclusive scan, task visiting order.
#pragma once
#include "common.cuh"
#include "msm_core.cuh"

namespace b381 {

// out[i] = sum of in[0..i) for i < n (in == out allowed); `copy`, when given, receives the same values (the scatter
// cursors).  `launches` counts the kernels launched.
cudaError_t exclusive_scan_u32(Scratch& sc, const uint32_t* in, uint32_t* out, size_t n, uint32_t* copy = nullptr,
                               int* launches = nullptr);

// pass 1: hist[0..nbuckets] (nbuckets + 1 entries, the last one stays 0) = pairs per bucket slot.
// host_src != nullptr: the scalars are still in (pinned or pageable) HOST memory; they are copied into d_scalars in
// chunks on a side stream and histogrammed as they land.
cudaError_t msm_histogram(Scratch& sc, const fr_t* d_scalars, bool scalars_mont, const msm_shape& sh, uint32_t* hist,
                          const fr_t* host_src = nullptr, int* launches = nullptr);

// scan + pass 2: offsets[0..nbuckets] = bucket boundaries (offsets[nbuckets] = n * W), vals[n * W] = entries
// (base index << 1 | sign) grouped by bucket slot, order inside a bucket unspecified.
cudaError_t msm_group_pairs(Scratch& sc, const fr_t* d_scalars, bool scalars_mont, const msm_shape& sh, const uint32_t* hist,
                            uint32_t* offsets, uint32_t* vals, int* launches = nullptr);

// order[0..ntasks) = task ids, longest task first (ties in unspecified order); K = task length bound
cudaError_t msm_task_order(Scratch& sc, size_t max_tasks, const uint32_t* ntasks_dev, const uint2* tasks, uint32_t K,
                           uint32_t* order, int* launches = nullptr);

}  // namespace b381
