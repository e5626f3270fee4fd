// K6a launcher.
#pragma once

#include "common.hpp"

namespace lvkb200
{

constexpr int RANSAC_HYPOTHESES = 256;
constexpr int RANSAC_REFINE_ITERS = 5;

struct RansacResult
{
    double h[9];
    int found;
    int inliers;
};

// All pointers are device memory.  d_models: HYP*9 floats, d_scores: HYP floats.
lvkb200_status ransac_homography(cudaStream_t cs, const float2* d_src, const float2* d_dst, int n, float threshold,
                                 float* d_models, float* d_scores, RansacResult* d_result, uint8_t* d_mask);

}  // namespace lvkb200
