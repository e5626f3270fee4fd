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

// fast_filter on the device: keeps a[i], b[i] where keep[i] != 0, in the reference's swap-erase order.
// d_perm / d_removed: scratch of n ints each.  *d_n_out receives the surviving count.
lvkb200_status compact_swap_erase(cudaStream_t cs, const float2* d_a, const float2* d_b, const uint8_t* d_keep,
                                  const TrackParams* d_params, float2* d_a_out, float2* d_b_out, int* d_perm,
                                  int* d_removed, int* d_n_out);

// Layout of the tracking chain's result block [matches float2 x cap | status u8 x cap | mask u8 x cap | pad | model]
// in device memory (`dev`) and its mirror in mapped pinned host memory (`host`, device view; nullptr = no copy).
// All offsets are multiples of 16 bytes.
struct TrackOutCopy
{
    const uint8_t* dev = nullptr;
    uint8_t* host = nullptr;
    uint32_t off_status = 0, off_mask = 0, off_result = 0;
};

// Copies the LK part of the result block (d_params->n matches + status) to the host mirror: local-motion mode, where
// no estimator kernel follows LK.
lvkb200_status track_out_copy(cudaStream_t cs, const TrackParams* d_params, const TrackOutCopy& out);

// d_* are device memory (the count too).  d_models: HYP*9 floats, d_scores: HYP + 4 floats
// (the last 8 bytes, 8-byte aligned, hold the packed arg-min key of the scoring pass).  With out.host set, the last
// kernel also copies the used part of the result block (d_params->n matches + status, mask, model) to the host mirror.
lvkb200_status ransac_homography(cudaStream_t cs, const float2* d_src, const float2* d_dst, const int* d_n,
                                 const TrackParams* d_params, float* d_models, float* d_scores,
                                 RansacResult* d_result, uint8_t* d_mask, const TrackOutCopy& out = TrackOutCopy{});

}  // namespace lvkb200
