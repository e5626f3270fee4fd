// K0 launcher state: per-geometry INTER_AREA tables resident on the device.
#pragma once

#include "common.hpp"
#include "stream.hpp"

namespace lvkb200
{

struct IngestPlan
{
    int sw = 0, sh = 0, dw = 0, dh = 0;
    int xcount = 0, ycount = 0;  // table entries per destination column / row
    int fast = 0;                // 0 = weighted tables, 1 = integer block mean, 2 = OpenCV's 2x2 special case
    float fast_scale = 0.f;
    int isx = 0, isy = 0;        // the integer scale factors when fast != 0
    DeviceBuffer d_xtab, d_ytab, d_xw, d_yw;

    // (Re)builds the tables when the geometry changed.
    lvkb200_status prepare(int src_w, int src_h, int dst_w, int dst_h, cudaStream_t cs);
    // Reads the frame once, writes the dw x dh detection image.
    lvkb200_status launch(cudaStream_t cs, const uint8_t* src, size_t pitch, lvkb200_format format, uint8_t* dst,
                          size_t dst_pitch) const;
    void release();
};

}  // namespace lvkb200
