// K5/K6a launchers.
#pragma once

#include "common.hpp"

namespace lvkb200
{

constexpr int RANSAC_HYPOTHESES = 256;
constexpr int RANSAC_REFINE_ITERS = 3;

struct RansacResult
{
    double h[9];
    int found;
    int inliers;
    int n;  // number of correspondences the estimator saw (after the device-side compaction)
    int pad;
};

// Copies two small blocks with a kernel (sources/destinations may be mapped pinned host memory): keeps the tracking
// chain off the copy engines.  Pointers must be 16-byte aligned and the allocations padded to multiples of 16 bytes.
lvkb200_status zero_copy_transfer(cudaStream_t cs, const void* src0, void* dst0, size_t bytes0, const void* src1,
                                  void* dst1, size_t bytes1);

// fast_filter on the device: keeps a[i], b[i] where keep[i] != 0, in the reference's swap-erase order.
// d_perm / d_removed: scratch of n ints each.  *d_n_out receives the surviving count.
lvkb200_status compact_swap_erase(cudaStream_t cs, const float2* d_a, const float2* d_b, const uint8_t* d_keep,
                                  const TrackParams* d_params, float2* d_a_out, float2* d_b_out, int* d_perm,
                                  int* d_removed, int* d_n_out);

// All pointers are device memory (the count too).  d_models: HYP*9 floats, d_scores: HYP floats.
lvkb200_status ransac_homography(cudaStream_t cs, const float2* d_src, const float2* d_dst, const int* d_n,
                                 const TrackParams* d_params, float* d_models, float* d_scores,
                                 RansacResult* d_result, uint8_t* d_mask);

}  // namespace lvkb200
