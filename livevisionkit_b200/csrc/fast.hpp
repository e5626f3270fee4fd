// K2 launcher state.
#pragma once

#include <vector>

#include "common.hpp"
#include "stream.hpp"

namespace lvkb200
{

constexpr int FAST_MAX_REGIONS = 16;

struct FastRegion
{
    int x, y, w, h;  // sub-image rectangle (cv::Rect conversion of FASTRegion::bounds)
    int threshold;
};

struct FastPoint
{
    short x, y;  // region-local pixel
    int score;   // cv::KeyPoint::response (integer valued)
};

struct FastDetector
{
    int w = 0, h = 0;
    size_t score_pitch = 0;
    int row_cap = 0, max_rows = 0, out_cap = 0, launched = 0;
    DeviceBuffer d_score, d_row_x, d_row_s, d_row_count;
    // The gather kernel writes the (few hundred) keypoints and the per-region counts straight into mapped pinned host
    // memory: no copy-engine transfer, and the host waits on `done` only, so work queued behind FAST keeps running.
    PinnedBuffer h_count, h_out;
    cudaEvent_t done = nullptr;

    lvkb200_status prepare(int width, int height);
    // Enqueues score -> NMS -> ordered gather for `n` regions of the device image.
    lvkb200_status launch(cudaStream_t cs, const uint8_t* img, size_t pitch, const FastRegion* regions, int n);
    // Waits for the last launch (not for the stream) and returns the per-region keypoint lists (OpenCV emission order).
    lvkb200_status fetch(std::vector<std::vector<FastPoint>>& out);
    void release();
};

}  // namespace lvkb200
