// K1/K4 launcher state: padded image pyramid + Scharr derivative planes of one detection image.
#pragma once

#include "common.hpp"
#include "stream.hpp"

namespace lvkb200
{

constexpr int LK_PAD = 11;            // winSize: every level carries an 11-px border
constexpr int LK_REF_MAX_LEVEL = 3;   // OPTICAL_TRACKER_PYR_LEVELS (FrameTracker.cpp:34)
constexpr int LK_MAX_LEVELS = LK_REF_MAX_LEVEL + 1;

struct LkPyramid
{
    int levels = 0;
    bool valid = false;  // built for the current frame
    int w[LK_MAX_LEVELS] = {}, h[LK_MAX_LEVELS] = {};
    size_t img_pitch[LK_MAX_LEVELS] = {}, deriv_pitch[LK_MAX_LEVELS] = {};
    DeviceBuffer img[LK_MAX_LEVELS], deriv[LK_MAX_LEVELS];

    lvkb200_status prepare(int width, int height);
    // det: device detection image (width x height, det_pitch).  Builds all levels and derivative planes.
    lvkb200_status build(cudaStream_t cs, const uint8_t* det, size_t det_pitch);
    void release();
};

// Squared stopping epsilon of the (call_index+1)-th calc() on one cv::SparsePyrLKOpticalFlow object (see lk.cu).
double lk_epsilon_for_call(int call_index);

// Tracks n points from `prev` to `next` (device arrays of float2 / uint8).
// The launch covers max_points; the frame's real count and stopping epsilon are read from d_params (device).
lvkb200_status lk_track(cudaStream_t cs, const LkPyramid& prev, const LkPyramid& next, const float2* d_prev_pts,
                        int max_points, const TrackParams* d_params, float2* d_next_pts, uint8_t* d_status);

}  // namespace lvkb200
