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

    lvkb200_status prepare(int width, int height, cudaStream_t cs);
    // det: device detection image (width x height, det_pitch).  Builds all levels and derivative planes.
    lvkb200_status build(cudaStream_t cs, const uint8_t* det, size_t det_pitch);
    void release();
};

// Squared stopping epsilon of the (call_index+1)-th calc() on one cv::SparsePyrLKOpticalFlow object (see lk.cu).
double lk_epsilon_for_call(int call_index);

// Inputs and outputs of one LK launch.  Every array holds max_points entries, max_points a multiple of 4, 16-byte
// aligned.  pts_in / prm_in may be device memory or the device view of MAPPED PINNED HOST memory (the kernel reads
// them itself: no transfer step before it); *_copy (optional) receive device copies for the kernels that follow;
// *_host (optional, device views of mapped pinned host memory) receive a second copy of the results.
struct LkIo
{
    const float2* pts_in = nullptr;
    const TrackParams* prm_in = nullptr;
    float2* pts_copy = nullptr;
    TrackParams* prm_copy = nullptr;
    float2* next = nullptr;
    uint8_t* status = nullptr;
    float2* next_host = nullptr;
    uint8_t* status_host = nullptr;
};

// The frame's parameters and points as ONE kernel argument (see k_lk_track): 51 x 51 suppression-grid cells, the
// largest feature capacity of the presets the reference ships, rounded up to the stream's allocation granule (64).
constexpr int LK_INLINE_POINTS = 2624;
struct LkPack
{
    TrackParams prm;
    int inline_points;  // 0: read io.prm_in / io.pts_in instead
    int pad;
    float2 pts[LK_INLINE_POINTS];
};

// Tracks n points from `prev` to `next`.  The launch covers max_points (a multiple of 4); the frame's real count and
// stopping epsilon come from pack->prm (pack != nullptr: inputs ride the launch as kernel arguments, max_points <=
// LK_INLINE_POINTS) or from io.prm_in / io.pts_in (device memory).
lvkb200_status lk_track(cudaStream_t cs, const LkPyramid& prev, const LkPyramid& next, int max_points, const LkIo& io,
                        const LkPack* pack = nullptr);

}  // namespace lvkb200
